"""Profiling target for ncu (GPU box): renders `steps` iterations of the bench workload, nothing else."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime
from ignis_b200.scene import load_scene

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scene = sys.argv[2] if len(sys.argv) > 2 else "diamond_scene.json"
t = load_scene(os.path.join(ROOT, "scenes", scene), 1920, 1080)
with Runtime(t, 1920, 1080, spi=4) as rt:
    if os.environ.get("DEFER0"):   # every path ends inside its own launch: the split-turn launches of the capture are exactly `steps` iterations' worth
        rt.device.setOption("defer_permille", 0)
    for _ in range(steps):
        rt.step()
    st = rt.device.getStatistics()
    work = rt.device.launchProfile()["k_turn_trace_work"]
print(st)
print("k_turn_trace work (rays traced by the split-turn trace launches):", work)
