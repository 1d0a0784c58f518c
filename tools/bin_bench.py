"""GPU box: material binning (a9) on / off on the C5 scene and on the textures-and-maps test scene: per-kernel times (CUDA events)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "aids")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
from debug_maps import build

def run(t, w, h, spi, label, opts):
    with Runtime(t, w, h, spi=spi) as rt:
        for k, v in opts.items():
            rt.device.setOption(k, v)
        for _ in range(3):
            rt.step()
        rt.reset(); rt.device.resetStatistics(); rt.device.setOption("profile_kernels", 1)
        n = 8
        for _ in range(n):
            rt.step()
        rt.device.sync()
        prof = rt.device.launchProfile()["kernels"]; st = rt.device.getStatistics()
        tot = sum(v["ms"] for v in prof.values())
        print(f"{label:28s} {opts} ms/step {tot / n:.3f}  shade {prof['k_turn_shade']['ms'] / n:.3f}  trace {prof['k_turn_trace']['ms'] / n:.3f}  wave {prof['k_wavefront']['ms'] / n:.3f}  Mrays/s {st['TotalRays'] / tot / 1e3:.0f}", flush=True)

if __name__ == "__main__":
    c5 = load_scene(os.path.join(ROOT, "scenes", "many_point_lights.json"))
    maps = load_scene(build(tempfile.mkdtemp())[0])
    for label, t, w, h, spi in (("many_point_lights 3840x2160", c5, 3840, 2160, 1), ("textures+maps 1920x1080", maps, 1920, 1080, 4)):
        for opts in ({"bin_materials": 0}, {"bin_materials": 1}, {"bin_materials": 0, "turn_shade_blocks": 2}, {"bin_materials": 1, "turn_shade_blocks": 2}):
            run(t, w, h, spi, label, opts)
