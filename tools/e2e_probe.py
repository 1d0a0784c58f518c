"""GPU box: where the end-to-end step (render + framebuffer read) spends its time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene

t = load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json"), 1920, 1080)
with Runtime(t, 1920, 1080, spi=4) as rt:
    dev = rt.device
    for k, v in (a.split("=") for a in sys.argv[1:]):
        dev.setOption(k, int(v))
    def run(mode, n=16):
        rt.reset(); rt.step(); dev.sync(); rt.reset()
        t0 = time.perf_counter()
        for _ in range(n):
            rt.step()
            if mode == "sync": dev.sync()
            elif mode == "read": rt.getFramebufferForHost()
        dev.sync()
        return (time.perf_counter() - t0) * 1e3 / n
    for mode in ("none", "sync", "read", "none", "sync", "read"):
        print(mode, f"{run(mode):.3f} ms/step", flush=True)
