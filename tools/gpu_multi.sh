#!/bin/bash
# One multi-GPU box call (gpurun --gpus 8): the headline workload at N = 8 / 4 / 2 as the driver launches it, C4 at N = 4, C5 at N = 8. Arg: tag
TAG=${1:-x}
mkdir -p gpurun_out
run() {  # n workload steps warmup extra-env
  local n=$1 wl=$2 k=$3 w=$4
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $n --workload $wl --steps $k --warmup $w 2> gpurun_out/multi_${TAG}_${wl}_n$n.err | tail -1 > gpurun_out/multi_${TAG}_${wl}_n$n.json
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/multi_${TAG}_${wl}_n$n.json"))
    print("$wl N=$n", round(d["value"]), "Mrays/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"]), " parity", d.get("parity", {}).get("rel_l2"), d.get("parity", {}).get("ok"))
    for r in d["per_rank_ms"]["ranks"]: print("   ", [round(x, 2) for x in r[:4]], r[5:])
except Exception as e:
    print("$wl N=$n FAILED", e)
P
}
if [ "$2" = "quick" ]; then run 8 c2 20 5; run 8 c5 16 3; exit 0; fi
if [ "$2" = "c2" ]; then run 8 c2 20 5; run 4 c2 20 5; exit 0; fi
if [ "$2" = "c2n8" ]; then run 8 c2 20 5; exit 0; fi
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,P2P run 8 c2 20 5
grep -E "via P2P|via SHM|NVLS|Connected|channels" gpurun_out/multi_${TAG}_c2_n8.err | sort | uniq -c | sort -rn | head -30 > gpurun_out/multi_${TAG}_nccl_transport_n8.txt
run 8 c5 16 3
run 4 c4 8 3
run 4 c2 20 5
run 2 c2 20 5
