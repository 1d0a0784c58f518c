#!/bin/bash
# GPU box: ncu --set full of the trace and shade kernels of one iteration; the reports are turned into the raw-page CSV and the per-line summaries
# ON the box (the .ncu-rep files are too large for gpurun_out). Arg: tag
TAG=${1:-x}
mkdir -p gpurun_out
for K in trace shade; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_turn_$K -c 3 -f -o /tmp/prof_$K python tools/profile_step.py 1 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_turn_${K}_full.csv 2>/dev/null
  ncu -i /tmp/prof_$K.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$K.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/src_$K.csv 45 > gpurun_out/${TAG}_k_turn_${K}_lines.txt 2>&1
  python tools/ncu_stalls.py /tmp/src_$K.csv long_sb 20 > gpurun_out/${TAG}_k_turn_${K}_stalls.txt 2>&1
  rm -f /tmp/prof_$K.ncu-rep /tmp/src_$K.csv
done
ls -la gpurun_out | grep $TAG
