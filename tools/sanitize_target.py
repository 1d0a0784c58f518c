"""GPU box: small workload for compute-sanitizer (memcheck / racecheck / synccheck): renders, drains, trace hooks."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.device import Runtime, RAY_DTYPE
from ignis_b200.scene import load_scene
for scene, w, h in (("diamond_scene.json", 96, 54), ("primitives.json", 64, 36), ("evaluation/multilight-hierarchy.json", 48, 48), ("evaluation/sun-on-plane.json", 48, 48)):
    t = load_scene(os.path.join(ROOT, "scenes", scene))
    with Runtime(t, w, h, spi=2) as rt:
        for split in (0, 2):
            rt.device.setOption("split_turns", split)
            rt.step(); rt.step()
            img = rt.getFramebufferForHost().copy()
        rng = np.random.default_rng(0)
        rays = np.zeros(3000, RAY_DTYPE)
        rays["org"] = rng.uniform(t.bbox_min, t.bbox_max, (3000, 3)); d = rng.normal(size=(3000, 3)); rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
        rays["tmin"], rays["tmax"] = 1e-3, 1e30
        for wide in (0, 1 << 20):
            rt.device.setOption("wide_rays_per_group", wide)
            hits = rt.device.traceClosest(rays); occ = rt.device.traceAny(rays)
        print(scene, float(img.mean()), int((hits["prim_id"] >= 0).sum()), int(occ.sum()))
