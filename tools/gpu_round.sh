#!/bin/bash
# One GPU-box call: parity tests, bench, ncu launch list + one full capture of the persistent kernel. Args: tag [full]
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc >> gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 16 --warmup 3 2> gpurun_out/bench_$TAG.err | tail -1 | tee gpurun_out/bench_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 2 > gpurun_out/launches_$TAG.log 2>&1
if [ "$2" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_turn_trace -c 3 -f -o gpurun_out/prof_$TAG python tools/profile_step.py 1 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_turn_shade -c 3 -f -o gpurun_out/prof_shade_$TAG python tools/profile_step.py 1 > gpurun_out/prof_shade_$TAG.log 2>&1
fi
