"""Debug helper (GPU box): prints the closest-hit records where the CUDA path and the oracle disagree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ignis_b200.device import B200Device, RAY_DTYPE
from ignis_b200.scene import load_scene
from oracle.oracle import Oracle
from test_gpu_parity import camera_rays

scene = sys.argv[1] if len(sys.argv) > 1 else "primitives.json"
t = load_scene(os.path.join(ROOT, "scenes", scene))
rays = camera_rays(t, 320, 180)
rng = np.random.default_rng(3)
extra = np.zeros(20000, RAY_DTYPE)
extra["org"] = rng.uniform(t.bbox_min, t.bbox_max, (20000, 3))
d = rng.normal(size=(20000, 3))
extra["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
extra["tmin"], extra["tmax"] = 1e-3, 3.4e38
rays = np.concatenate([rays, extra])
o = Oracle(t)
ref = o.trace_closest(rays, use_bvh=True)
brute = o.trace_closest(rays, use_bvh=False)
with B200Device() as dev:
    dev.assignScene(t)
    got = dev.traceClosest(rays)
bad = np.nonzero((got["ent_id"] != ref["ent_id"]) | (got["prim_id"] != ref["prim_id"]) | (got["t"].view(np.uint32) != ref["t"].view(np.uint32)))[0]
print("rays", len(rays), "mismatches", len(bad), "oracle bvh vs brute mismatches", int(((ref["ent_id"] != brute["ent_id"]) | (ref["prim_id"] != brute["prim_id"])).sum()))
print("shape types", t.shape_lookups["type_id"], "entity shapes", t.entities[:, 33].view(np.int32))
for i in bad[:25]:
    print(i, "ray", rays[i], "\n   gpu", got[i], "ref", ref[i], "brute", brute[i])
from collections import Counter
print("ref ents of mismatches", Counter(ref["ent_id"][bad].tolist()), "gpu ents", Counter(got["ent_id"][bad].tolist()))
