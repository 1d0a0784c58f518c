#!/bin/bash
# gpurun --gpus 2: everything that needs two GPUs in one process or two ranks (C-ABI exchange tests, the multi-GPU plugin device)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_plugin_host.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_two_${1:-x}.log
