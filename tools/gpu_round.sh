#!/bin/bash
# One GPU-box call: parity tests, bench, ncu launch list + full capture of the traversal kernel. Args: tag
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 16 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 1 > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${2:-k_trace_primary}" -c 2 -f -o gpurun_out/prof_$TAG python tools/profile_step.py 1 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
