import os, sys
sys.path.insert(0, "/root/repo")
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
for scene in ("diamond_scene.json", "primitives.json"):
    t = load_scene(os.path.join("/root/repo/scenes", scene))
    for mode in (0, 1, 2):
        with Runtime(t, 1920, 1080, spi=4) as rt:
            rt.device.setOption("bin_materials", mode)
            rt.device.assignScene(t)   # classes are assigned at scene upload
            for _ in range(3): rt.step()
            rt.reset(); rt.device.resetStatistics(); rt.device.setOption("profile_kernels", 1)
            n = 16
            for _ in range(n): rt.step()
            rt.device.sync()
            prof = rt.device.launchProfile()["kernels"]; st = rt.device.getStatistics()
            tot = sum(v["ms"] for v in prof.values())
            print(scene, "bin", mode, f"ms/step {tot/n:.3f} shade {prof['k_turn_shade']['ms']/n:.3f} trace {prof['k_turn_trace']['ms']/n:.3f} wave {prof['k_wavefront']['ms']/n:.3f} Mrays/s {st['TotalRays']/tot/1e3:.0f}", flush=True)
