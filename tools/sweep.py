"""GPU box: sweeps of option COMBINATIONS on full-size workloads (per-kernel CUDA-event times) with a frame check against the first combination.
Usage: sweep.py "k=v,k=v;k=v;..." [scene[:WxH[:spi]] ...]   (an empty combination = the defaults)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene

combos = []
for part in (sys.argv[1] if len(sys.argv) > 1 else "").split(";"):
    combos.append({k: int(v) for k, v in (kv.split("=") for kv in part.split(",") if kv)})
for spec in sys.argv[2:] or ["diamond_scene.json"]:
    f = spec.split(":")
    scene = f[0]
    w, h = (int(x) for x in f[1].split("x")) if len(f) > 1 else (1920, 1080)
    spi = int(f[2]) if len(f) > 2 else 4
    t = load_scene(os.path.join(ROOT, "scenes", scene), w, h)
    ref = None
    for opts in combos:
        with Runtime(t, w, h, spi=spi) as rt:
            for k, v in opts.items():
                rt.device.setOption(k, v)
            rt.device.assignScene(t)   # some options act when the scene is uploaded
            for _ in range(3): rt.step()
            rt.reset(); rt.device.resetStatistics(); rt.device.setOption("profile_kernels", 1)
            n = 16
            for _ in range(n): rt.step()
            rt.device.sync()
            prof = rt.device.launchProfile()["kernels"]; st = rt.device.getStatistics()
            img = rt.getFramebufferForHost().copy()
            tot = sum(x["ms"] for x in prof.values())
            if ref is None: ref = (img, st["TotalRays"])
            err = float(np.linalg.norm((img - ref[0]).ravel()) / max(np.linalg.norm(ref[0].ravel()), 1e-30))
            print(f"{scene} {w}x{h} spi {spi} {opts} ms/step {tot/n:.3f} shade {prof['k_turn_shade']['ms']/n:.3f} trace {prof['k_turn_trace']['ms']/n:.3f} "
                  f"wave {prof['k_wavefront']['ms']/n:.3f} Mrays/s {st['TotalRays']/tot/1e3:.0f} | vs first: rel_l2 {err:.2e} rays {'same' if st['TotalRays'] == ref[1] else st['TotalRays'] - ref[1]}", flush=True)
