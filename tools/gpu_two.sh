#!/bin/bash
# gpurun --gpus 2: everything that needs two GPUs in one process or two ranks (C-ABI exchange tests, the multi-GPU plugin device), then the bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_plugin_host.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_two_${1:-x}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/two_${1:-x}_n2.err | tail -1 > gpurun_out/two_${1:-x}_n2.json
python - <<P
import json
d = json.load(open("gpurun_out/two_${1:-x}_n2.json"))
print("N=2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["rel_l2"], [[round(x, 2) for x in r] for r in d["per_rank_ms"]["ranks"]])
P
tail -3 gpurun_out/two_${1:-x}_n2.err
