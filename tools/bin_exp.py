"""GPU box: option sweeps on the full-size workloads (per-kernel CUDA-event times). Usage: bin_exp.py option v1,v2,.. [scene ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
opt = sys.argv[1] if len(sys.argv) > 1 else "bin_materials"
vals = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2").split(",")]
scenes = sys.argv[3:] or ["diamond_scene.json", "primitives.json"]
for scene in scenes:
    t = load_scene(os.path.join(ROOT, "scenes", scene))
    for v in vals:
        with Runtime(t, 1920, 1080, spi=4) as rt:
            rt.device.setOption(opt, v)
            rt.device.assignScene(t)   # some options act when the scene is uploaded
            for _ in range(3): rt.step()
            rt.reset(); rt.device.resetStatistics(); rt.device.setOption("profile_kernels", 1)
            n = 16
            for _ in range(n): rt.step()
            rt.device.sync()
            prof = rt.device.launchProfile()["kernels"]; st = rt.device.getStatistics()
            tot = sum(x["ms"] for x in prof.values())
            print(scene, opt, v, f"ms/step {tot/n:.3f} shade {prof['k_turn_shade']['ms']/n:.3f} trace {prof['k_turn_trace']['ms']/n:.3f} wave {prof['k_wavefront']['ms']/n:.3f} Mrays/s {st['TotalRays']/tot/1e3:.0f}", flush=True)
