"""GPU box: stand-alone trace-phase throughput (igb200_bench_trace) for occupancy variants of k_trace."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import B200Device, RAY_DTYPE
from ignis_b200.scene import load_scene

scene = sys.argv[1] if len(sys.argv) > 1 else "diamond_scene.json"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
t = load_scene(os.path.join(ROOT, "scenes", scene), 1920, 1080)
rng = np.random.default_rng(0)
rays = np.zeros(n, RAY_DTYPE)
rays["org"] = rng.uniform(t.bbox_min * 0.98, t.bbox_max * 0.98, (n, 3))
d = rng.normal(size=(n, 3)); rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
rays["tmin"], rays["tmax"] = 1e-3, 3.0e38
with B200Device() as dev:
    dev.assignScene(t)
    for budget in (40960, 20480):
        dev.setOption("stage_budget", budget)
        for tb in (2, 3, 4):
            if tb == 4 and budget > 20480:
                continue
            dev.setOption("trace_blocks", tb)
            for wide in (0,):
                dev.setOption("wide_rays_per_group", wide)
                ms_c = dev.benchTrace(rays, any_hit=False, repeat=5)
                ms_a = dev.benchTrace(rays, any_hit=True, repeat=5)
                print(f"stage {budget} trace_blocks {tb}: closest {n / ms_c / 1e6:.2f} Grays/s ({ms_c:.3f} ms)  any {n / ms_a / 1e6:.2f} Grays/s ({ms_a:.3f} ms)", flush=True)
