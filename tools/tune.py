"""GPU box: sweeps the persistent-kernel options on the bench workload and prints ms per step for each."""
import itertools, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene

scene = sys.argv[1] if len(sys.argv) > 1 else "diamond_scene.json"
opts = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    opts[k] = [int(x) for x in v.split(",")]
keys = list(opts)
t = load_scene(os.path.join(ROOT, "scenes", scene), 1920, 1080)
with Runtime(t, 1920, 1080, spi=4) as rt:
    if os.environ.get("PART"):   # emulate one rank of a multi-GPU run: PART=rank,world
        r_, w_ = (int(x) for x in os.environ["PART"].split(","))
        rt.device.setPartition(r_, w_, 32)
    for combo in itertools.product(*[opts[k] for k in keys]):
        for k, v in zip(keys, combo):
            rt.device.setOption(k, v)
        rt.reset(); rt.step()
        rt.reset(); rt.device.resetStatistics()
        n = 8
        import time as _t
        t0 = _t.perf_counter()
        for _ in range(n):
            rt.step()
        rt.device.sync(); wall = (_t.perf_counter() - t0) * 1e3 / n
        st = rt.device.getStatistics(); kt = rt.device.kernelTimes()
        print(dict(zip(keys, combo)), f"wall/step {wall:.3f} ms/step {st['render_ms'] / n:.3f}  Mrays/s {st['TotalRays'] / st['render_ms'] / 1e3:.0f}  trace {kt['trace']['ms'] / n:.3f} ms  shade {kt['shade_generate']['ms'] / n:.3f} ms  phases {kt['trace']['launches'] // n}", flush=True)
    if os.environ.get("TURNLOG"):
        items, tr, sh = rt.device.turnLog()
        print("turn items trace_us shade_us")
        for k in range(len(items)):
            print(k, items[k], round(tr[k] / 1e3, 1), round(sh[k] / 1e3, 1))
        print("sum trace ms", tr.sum() / 1e6, "shade ms", sh.sum() / 1e6)
    if os.environ.get("IGB200_STEP_STATS") or os.environ.get("IGB200_PROBE"):
        print(rt.device.stepStats())
