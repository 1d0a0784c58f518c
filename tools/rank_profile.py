"""GPU box: per-kernel CUDA-event times of one rank's share of the bench frame (PART=rank,world emulates a multi-GPU rank)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene
r_, w_ = (int(x) for x in os.environ.get("PART", "0,1").split(","))
t = load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json"), 1920, 1080)
with Runtime(t, 1920, 1080, spi=4) as rt:
    rt.device.setPartition(r_, w_, 32)
    for a in sys.argv[1:]:
        k, v = a.split("="); rt.device.setOption(k, int(v))
    for prof in (0, 1):
        rt.reset()
        for _ in range(3):
            rt.step()
        rt.reset(); rt.device.resetStatistics()
        rt.device.setOption("profile_kernels", prof)
        n = 16
        t0 = time.perf_counter()
        for _ in range(n):
            rt.step()
        rt.device.sync(); wall = (time.perf_counter() - t0) * 1e3 / n
        st = rt.device.getStatistics()
        print(f"profile={prof} wall/step {wall:.3f} ms, device kernel time/step {st['render_ms'] / n:.3f} ms, rays/step {st['TotalRays'] / n:.0f}")
        if prof:
            p = rt.device.launchProfile()
            tot = 0
            for k, v in p["kernels"].items():
                print(f"   {k:14s} {v['ms'] / n:.3f} ms/step in {v['launches'] / n:.2f} launches ({1e3 * v['ms'] / max(v['launches'], 1):.1f} us each)")
                tot += v["ms"] / n
            print(f"   sum {tot:.3f} ms/step")
