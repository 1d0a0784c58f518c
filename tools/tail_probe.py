"""GPU box: turn log of tiny frames (pure latency / fixed overhead of a wavefront turn)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
for (w, h, spi) in [(64, 36, 1), (256, 144, 1), (640, 360, 1)]:
    t = load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json"), w, h)
    with Runtime(t, w, h, spi=spi) as rt:
        for a in sys.argv[1:]:
            k, v = a.split("="); rt.device.setOption(k, int(v))
        rt.step(); rt.reset(); rt.device.resetStatistics(); rt.step()
        items, tr, sh = rt.device.turnLog()
        st = rt.device.getStatistics()
        print(w, h, spi, "ms", st["render_ms"], "turns", len(items))
        print(" items", list(items[:12]), "...", list(items[-3:]))
        print(" trace_us", [round(x / 1e3, 1) for x in tr[:12]], "...", [round(x / 1e3, 1) for x in tr[-3:]])
        print(" shade_us", [round(x / 1e3, 1) for x in sh[:12]], "...", [round(x / 1e3, 1) for x in sh[-3:]])
