"""GPU box: time of igb200_set_scene (host BVH8 build + upload) for the large synthetic scene."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import B200Device
from ignis_b200.scene import load_scene
for name in ("diamond_scene.json", "primitives.json", "synthetic_room.json"):
    t0 = time.perf_counter(); t = load_scene(os.path.join(ROOT, "scenes", name)); t1 = time.perf_counter()
    with B200Device() as dev:
        t2 = time.perf_counter(); dev.assignScene(t); t3 = time.perf_counter()
    print(f"{name}: ingest {t1 - t0:.2f} s, assignScene (BVH8 build + upload) {t3 - t2:.3f} s, unique triangles in shapes: {t.shape_data.nbytes / 1e6:.1f} MB blob, instanced {t.n_triangles}")
