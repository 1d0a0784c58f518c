"""GPU box: renders the textures-and-maps test scene one material at a time against the oracle (debug aid)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import furnace_scene
from ignis_b200.scene import load_scene
from oracle.oracle import Oracle

def build(tmp, only=None):
    rng = np.random.default_rng(8)
    np.save(os.path.join(tmp, "a.npy"), rng.random((9, 13, 3)).astype(np.float32))
    np.save(os.path.join(tmp, "n.npy"), (0.5 + 0.5 * np.stack([0.3 * rng.standard_normal((16, 16)), 0.3 * rng.standard_normal((16, 16)), np.ones((16, 16))], axis=2) / np.sqrt(1.18)).astype(np.float32).clip(0, 1))
    a, n = os.path.join(tmp, "a.npy"), os.path.join(tmp, "n.npy")
    tex = [{"type": "checkerboard", "name": "check", "scale_x": 6, "scale_y": 3, "color0": [0.2, 0.3, 0.4], "color1": [0.9, 0.8, 0.7], "transform": {"rotate": [0, 0, 30]}},
           {"type": "image", "name": "near", "filename": a, "filter_type": "nearest", "wrap_mode": "mirror"},
           {"type": "image", "name": "lin", "filename": a, "filter_type": "bilinear", "wrap_mode_u": "clamp", "wrap_mode_v": "repeat", "transform": {"scale": [2.5, 1.5, 1]}},
           {"type": "image", "name": "cub", "filename": a, "transform": [1.7, 0.2, -0.3, -0.1, 2.2, 0.4, 0, 0, 1]},
           {"type": "bitmap", "name": "bump", "filename": os.path.join(ROOT, "scenes", "textures/bumpmap.png")},
           {"type": "image", "name": "nrm", "filename": n, "filter_type": "bilinear"}]
    s = furnace_scene()
    s["technique"]["max_depth"] = 8
    s["textures"] = tex
    s["bsdfs"] = [{"type": "diffuse", "name": "d_check", "reflectance": "check"}, {"type": "diffuse", "name": "d_near", "reflectance": "near"},
                  {"type": "diffuse", "name": "d_lin", "reflectance": "lin"}, {"type": "conductor", "name": "c_cub", "specular_reflectance": "cub", "roughness": 0.3},
                  {"type": "dielectric", "name": "g_tex", "specular_reflectance": "lin", "specular_transmittance": "cub"},
                  {"type": "conductor", "name": "rc", "roughness": 0.2, "material": "copper"},
                  {"type": "bumpmap", "name": "b_rc", "bsdf": "rc", "map": "bump", "strength": 0.35},
                  {"type": "normalmap", "name": "n_d", "bsdf": "d_check", "map": "nrm", "strength": 0.8},
                  {"type": "normalmap", "name": "n_rc", "bsdf": "rc", "map": "nrm"}]
    s["shapes"] = [{"type": "cube", "name": "Box", "width": 1.0, "height": 1.0, "depth": 1.0, "origin": [-0.5, -0.5, -0.5]},
                   {"type": "rectangle", "name": "Floor", "width": 12, "height": 12, "origin": [-6, -6, -0.9]},
                   {"type": "sphere", "name": "Ball", "radius": 0.45}, {"type": "uvsphere", "name": "UV", "radius": 0.45}]
    names = ["d_near", "d_lin", "c_cub", "g_tex", "b_rc", "n_d", "n_rc"]
    s["entities"] = [{"name": "Floor", "shape": "Floor", "bsdf": "d_check"}]
    for k, b in enumerate(names):
        if only is not None and k != only: continue
        # (no map on the analytic sphere: it is one-sided -- face_normal is not flipped towards rays that hit it from inside, and
        # ensure_valid_reflection, core/sampling.art:120-165, then normalises a zero vector: NaN in the reference as well)
        shape = {"d_near": "Box", "d_lin": "Ball", "c_cub": "Ball", "g_tex": "Box", "b_rc": "UV", "n_d": "Box", "n_rc": "UV"}[b]
        s["entities"].append({"name": f"e{k}", "shape": shape, "bsdf": b, "transform": [{"translate": [-2.4 + 0.8 * k, 0.6 * ((k % 2) * 2 - 1), 0]}, {"rotate": [10 * k, 20, 5 * k]}]})
    s["camera"]["transform"] = {"lookat": {"origin": [0.5, -6.5, 3.0], "target": [0, 0, 0], "up": [0, 0, 1]}}
    s["lights"] = [{"type": "env", "name": "env", "radiance": [0.6, 0.7, 0.8]}, {"type": "point", "name": "p", "position": [1, -2, 3], "intensity": [15, 14, 13]}]
    return s, names

if __name__ == "__main__":
    tmp = tempfile.mkdtemp()
    gpu = "--cpu" not in sys.argv
    if gpu:
        from ignis_b200.device import Runtime
    for only in [None] + list(range(7)):
        s, names = build(tmp, only)
        t = load_scene(s)
        w, h = 240, 160
        ref = np.zeros((h, w, 3), np.float32)
        o = Oracle(t)
        for it in range(2): o.render(w, h, spi=4, iteration=it, fb=ref)
        msg = f"{'all' if only is None else names[only]}: oracle finite {np.isfinite(ref).all()} mean {ref.mean():.4f}"
        if gpu:
            with Runtime(t, w, h, spi=4) as rt:
                rt.step(); rt.step()
                got = rt.getFramebufferForHost().copy(); st = rt.device.getStatistics()
            bad = ~np.isfinite(got).all(axis=2)
            err = np.linalg.norm((np.nan_to_num(got) - ref).ravel()) / np.linalg.norm(ref.ravel())
            msg += f" | gpu nonfinite px {int(bad.sum())} rel_l2 {err:.3e} counters {st['CameraRayCount'], st['ShadowRayCount'], st['BounceRayCount']} vs {tuple(int(x) for x in o.counters)}"
            if bad.any():
                ys, xs = np.where(bad); msg += f" first bad {ys[0], xs[0]}"
        print(msg, flush=True)
