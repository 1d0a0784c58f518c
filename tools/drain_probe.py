"""GPU box: what the drain of the deferred tail costs for one rank's share of the bench frame (PART=rank,world): CUDA-event time of the drain and the
turn log of its persistent-kernel launch (rays, trace / shade time per turn)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
r_, w_ = (int(x) for x in os.environ.get("PART", "0,8").split(","))
K = int(os.environ.get("K", "20"))
t = load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json"), 1920, 1080)
with Runtime(t, 1920, 1080, spi=4) as rt:
    rt.device.setPartition(r_, w_, 32)
    for a in sys.argv[1:]:
        k, v = a.split("="); rt.device.setOption(k, int(v))
    stream = torch.cuda.ExternalStream(rt.device.stream(), device=torch.device("cuda", 0))
    for rep in range(3):
        rt.reset(); rt.device.resetStatistics()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record(stream)
        for _ in range(K): rt.step()
        e[1].record(stream)
        rt.device.sync()
        e[2].record(stream); stream.synchronize()
        st = rt.device.getStatistics()
        print(f"PART={r_},{w_} K={K} {sys.argv[1:]}: issued {e[0].elapsed_time(e[1]):.2f} ms, flush + drain {e[1].elapsed_time(e[2]):.2f} ms, rays {st['TotalRays']}, launches {st['KernelLaunches']}")
    items, tr, sh = rt.device.turnLog()
    n = len(items)
    print(" drain k_wavefront turns", n, "trace", round(sum(tr) / 1e6, 3), "ms shade", round(sum(sh) / 1e6, 3), "ms")
    print(" items", list(items[:8]), "...", list(items[-4:]))
    print(" trace_us", [round(x / 1e3, 1) for x in tr[:8]], "...", [round(x / 1e3, 1) for x in tr[-4:]])
    print(" shade_us", [round(x / 1e3, 1) for x in sh[:8]], "...", [round(x / 1e3, 1) for x in sh[-4:]])
