"""GPU box: host (binned SAH) against GPU (LBVH) BVH8 construction on one large mesh -- build time, tree quality, trace speed.
Usage: bvh_bench.py [subdivisions=8]  (icosphere: 20 * 4^s faces)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene = json.load(open(os.path.join(ROOT, "scenes", "synthetic_room.json")))
scene["shapes"] = [s for s in scene["shapes"] if not s["name"].startswith("Ico")] + [{"type": "icosphere", "name": "Big", "radius": 2.5, "subdivisions": sub}]
scene["entities"] = [e for e in scene["entities"] if not e["name"].startswith("S")] + [{"name": "B", "shape": "Big", "bsdf": "red", "transform": {"translate": [0, 0, 2.6]}}]
t0 = time.perf_counter(); t = load_scene(scene); print(f"ingest {time.perf_counter() - t0:.2f} s, {t.n_triangles} triangles", flush=True)
for gpu in (0, 1):
    with Runtime(t, 1920, 1080, spi=4) as rt:
        rt.device.setOption("gpu_bvh", gpu)
        t0 = time.perf_counter(); rt.device.assignScene(t); dt = time.perf_counter() - t0
        info = rt.device.sceneBuildInfo()
        for _ in range(2): rt.step()
        rt.reset(); rt.device.resetStatistics(); rt.device.setOption("profile_kernels", 1)
        n = 8
        for _ in range(n): rt.step()
        rt.device.sync()
        prof = rt.device.launchProfile()["kernels"]; st = rt.device.getStatistics()
        tot = sum(x["ms"] for x in prof.values())
        print(f"gpu_bvh {gpu}: assignScene {dt*1e3:.0f} ms (trees {info['build_us']/1e3:.1f} ms, {info['nodes']} nodes, host {info['host_built']} gpu {info['gpu_built']})  "
              f"render ms/step {tot/n:.3f} trace {prof['k_turn_trace']['ms']/n:.3f} Mrays/s {st['TotalRays']/tot/1e3:.0f}", flush=True)
