"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars: hit records -- entity / primitive ids exact, t/u/v bit-exact; radiance -- relative L2 <= 1e-4 (north_star),
measured values are ~1e-7 (float atomics reorder additions inside one pixel); ray counters exact.
"""
import os

import numpy as np
import pytest

from conftest import flat_scene, furnace_scene
from ignis_b200.device import B200Device, RAY_DTYPE, Runtime
from ignis_b200.scene import load_scene
from oracle.oracle import Oracle, detmath

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL_L2_TOL = 1e-4


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def scene_path(name):
    return os.path.join(ROOT, "scenes", name)


def camera_rays(tables, w, h, seed=0):
    cam = tables.camera
    rng = np.random.default_rng(seed)
    eye = np.asarray(cam["eye"], np.float32)
    d, up = np.asarray(cam["dir"], np.float64), np.asarray(cam["up"], np.float64)
    right = np.cross(d, up)
    right /= np.linalg.norm(right)
    sx = np.tan(float(cam["fov"]) / 2)
    sy = sx / (w / h)
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    nx = 2 * (xs + rng.random(xs.shape)) / w - 1
    ny = 1 - 2 * (ys + rng.random(ys.shape)) / h
    dirs = right[None, None] * (sx * nx)[..., None] + up[None, None] * (sy * ny)[..., None] + d[None, None]
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    rays = np.zeros(w * h, RAY_DTYPE)
    rays["org"] = eye
    rays["dir"] = dirs.reshape(-1, 3)
    rays["tmin"], rays["tmax"] = float(cam["tmin"]), float(cam["tmax"])
    return rays


def test_detmath_bit_exact():
    rng = np.random.default_rng(1)
    with B200Device() as dev:
        x = rng.uniform(-20, 20, 1 << 18).astype(np.float32)
        for fn in ("sin", "cos"):
            np.testing.assert_array_equal(dev.testDetmath(fn, x), detmath(fn, x))
        a = np.concatenate([rng.uniform(-1, 1, 1 << 18), [-1, 1, 0, 0.5, -0.5]]).astype(np.float32)
        np.testing.assert_array_equal(dev.testDetmath("acos", a), detmath("acos", a))
        y = rng.uniform(-3, 3, 1 << 18).astype(np.float32)
        np.testing.assert_array_equal(dev.testDetmath("atan2", y, x), detmath("atan2", y, x))


@pytest.mark.parametrize("scene,w,h", [("single_triangle.json", 256, 256), ("diamond_scene.json", 320, 180), ("primitives.json", 320, 180),
                                       ("evaluation/cbox-d6.json", 128, 128), ("synthetic_room.json", 192, 108)])
def test_closest_hit_records_match_oracle(scene, w, h):
    t = load_scene(scene_path(scene))
    rays = camera_rays(t, w, h)
    # add incoherent rays from inside the scene bounds
    rng = np.random.default_rng(3)
    extra = np.zeros(20000, RAY_DTYPE)
    lo, hi = t.bbox_min, t.bbox_max
    extra["org"] = rng.uniform(lo, hi, (20000, 3))
    d = rng.normal(size=(20000, 3))
    extra["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    extra["tmin"], extra["tmax"] = 1e-3, 3.4e38
    rays = np.concatenate([rays, extra])
    o = Oracle(t)
    ref = o.trace_closest(rays, use_bvh=True)
    ref_occ = o.trace_any(rays, flags=np.full(len(rays), 8, np.uint32))
    # both walks of the device: one lane per ray (large queues) and eight lanes per ray (small queues, traverse.cuh trace_wide);
    # with the scene copy in shared memory when it fits (the default), with nothing staged, and with a staged prefix (nodes and
    # triangles partly in shared memory, partly read through L1/L2) -- the closest hit must not depend on any of it
    # ... nor on the merged single-level tree small scenes are walked with by default ("flat": 0 = the two-level walk)
    for wide, opts in ((0, {}), (0, {"flat": 0}), (1 << 20, {}), (1 << 20, {"flat": 0}), (0, {"specialise_where": 0}), (0, {"stage_budget": 0}), (1 << 20, {"stage_budget": 0}),
                       (0, {"stage_partial": 1, "stage_budget": 6144}), (1 << 20, {"stage_partial": 1, "stage_budget": 6144})):
        budget = tuple(opts.items())
        with B200Device() as dev:
            dev.assignScene(t)
            dev.setOption("wide_rays_per_group", wide)
            for k, v in opts.items():
                dev.setOption(k, v)
            got = dev.traceClosest(rays)
            occ = dev.traceAny(rays)
        assert (got["ent_id"] == ref["ent_id"]).all(), (wide, budget)
        assert (got["prim_id"] == ref["prim_id"]).all(), (wide, budget)
        for k in ("t", "u", "v"):
            np.testing.assert_array_equal(got[k].view(np.uint32), ref[k].view(np.uint32))
        np.testing.assert_array_equal(occ, ref_occ)
    assert (ref["prim_id"] >= 0).any()


def render_both(tables, w, h, spi, iters, seed=0):
    o = Oracle(tables)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(iters):
        o.render(w, h, spi=spi, iteration=it, seed=seed, fb=ref)
    with Runtime(tables, w, h, spi=spi, seed=seed) as rt:
        for _ in range(iters):
            rt.step()
        got = rt.getFramebufferForHost().copy()
        stats = rt.device.getStatistics()
    return got, ref, stats, o.counters


@pytest.mark.parametrize("scene,w,h,spi,iters", [
    ("single_triangle.json", 256, 256, 1, 1),            # BASELINE config C1
    ("diamond_scene.json", 240, 135, 4, 2),              # C2 at reduced size (same spi)
    ("primitives.json", 240, 135, 4, 1),                 # C3 at reduced size
    ("evaluation/cbox-d6.json", 128, 128, 2, 2),
    ("evaluation/multilight-uniform.json", 128, 128, 2, 1),
    ("evaluation/sun-on-plane.json", 128, 128, 2, 2),           # sun: infinite cone light (light/sun.art)
    ("evaluation/multilight-simple.json", 128, 128, 2, 2),      # flux-CDF light selector (light_selector.art:46-77)
    ("evaluation/multilight-hierarchy.json", 128, 128, 2, 2),   # light hierarchy (light/light_hierarchy.art)
    ("evaluation/emissive-plane.json", 128, 128, 1, 1),
    ("evaluation/point.json", 64, 64, 1, 1),
    ("evaluation/sphere-light-pure.json", 128, 128, 2, 2),
    ("evaluation/two-planes-mirror.json", 128, 128, 4, 2),
    ("evaluation/room.json", 128, 128, 4, 1),                 # OBJ mesh, constant light    # mirror (smooth conductor) + tiny sphere light   # analytic sphere area light (light/area.art:260-316)
    ("synthetic_room.json", 192, 108, 2, 2),             # stand-in for C4: 1.8 M instanced triangles, geometry read through L2
    ("many_point_lights.json", 200, 200, 2, 3),          # C5 at reduced size: hierarchy over 10 embedded point lights + sky (2-D cdf), checkerboard, bump-mapped rough conductor
    ("evaluation/env4k-conditional.json", 128, 128, 2, 2),   # textured environment through the 2-D cdf, bicubic filter
    ("evaluation/env4k-none.json", 128, 128, 2, 2),          # ... sampled uniformly (make_environment_light over a texture)
    ("evaluation/env.json", 128, 128, 2, 2),                 # 100 x 50 8-bit map, nearest filter, MIS-compensated cdf
])
def test_radiance_matches_oracle(scene, w, h, spi, iters):
    t = load_scene(scene_path(scene))
    got, ref, stats, cnt = render_both(t, w, h, spi, iters)
    assert np.isfinite(got).all()
    assert ref.sum() > 0
    assert rel_l2(got, ref) <= REL_L2_TOL
    # ray counters are discrete: camera, shadow, bounce must agree exactly
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in cnt)


def _check(scene, w, h, spi, iters):
    t = load_scene(scene)
    got, ref, stats, cnt = render_both(t, w, h, spi, iters)
    assert np.isfinite(got).all() and ref.sum() > 0
    assert rel_l2(got, ref) <= REL_L2_TOL
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in cnt)
    return t


@pytest.mark.parametrize("roughness", [0.16, 0.5, 0.00005, {"roughness_u": 0.05, "roughness_v": 0.45}, {"roughness": 0.3, "anisotropic": 0.6}])
def test_rough_conductor_matches_oracle(roughness):
    """GGX / VNDF conductor (bsdf/conductor.art:45-141): NEE with its eval / pdf, VNDF sampling, rejected samples; gold instead of the
    lossless default so that the Fresnel term matters; a roughness <= 1e-4 is the delta distribution."""
    s = furnace_scene()
    s["bsdfs"] = [dict({"type": "conductor", "name": "glass", "material": "gold"}, **(roughness if isinstance(roughness, dict) else {"roughness": roughness}))]
    s["lights"].append({"type": "point", "name": "p", "position": [3, -2, 4], "intensity": [20, 20, 20]})
    s["shapes"].append({"type": "rectangle", "name": "Floor", "width": 8, "height": 8, "origin": [-4, -4, -0.9]})
    s["bsdfs"].append({"type": "diffuse", "name": "floor", "reflectance": [0.7, 0.6, 0.5]})
    s["entities"].append({"name": "Floor", "shape": "Floor", "bsdf": "floor"})
    t = _check(s, 128, 128, 4, 2)
    assert int(t.materials[0]["distribution"]) == 1


def test_textures_and_maps_match_oracle(tmp_path):
    """Checkerboard and image textures (all three filters, all three borders, a uv transform) on diffuse, dielectric and conductor
    colour parameters; a bump map and a normal map over a rough conductor and a diffuse BSDF; textured sphere (tex_coords = prim_coords)."""
    rng = np.random.default_rng(8)
    np.save(tmp_path / "a.npy", rng.random((9, 13, 3)).astype(np.float32))
    np.save(tmp_path / "n.npy", (0.5 + 0.5 * np.stack([0.3 * rng.standard_normal((16, 16)), 0.3 * rng.standard_normal((16, 16)), np.ones((16, 16))], axis=2) /
                                 np.sqrt(1.18)).astype(np.float32).clip(0, 1))
    tex = [{"type": "checkerboard", "name": "check", "scale_x": 6, "scale_y": 3, "color0": [0.2, 0.3, 0.4], "color1": [0.9, 0.8, 0.7], "transform": {"rotate": [0, 0, 30]}},
           {"type": "image", "name": "near", "filename": str(tmp_path / "a.npy"), "filter_type": "nearest", "wrap_mode": "mirror"},
           {"type": "image", "name": "lin", "filename": str(tmp_path / "a.npy"), "filter_type": "bilinear", "wrap_mode_u": "clamp", "wrap_mode_v": "repeat",
            "transform": {"scale": [2.5, 1.5, 1]}},
           {"type": "image", "name": "cub", "filename": str(tmp_path / "a.npy"), "transform": [1.7, 0.2, -0.3, -0.1, 2.2, 0.4, 0, 0, 1]},
           {"type": "bitmap", "name": "bump", "filename": scene_path("textures/bumpmap.png")},
           {"type": "image", "name": "nrm", "filename": str(tmp_path / "n.npy"), "filter_type": "bilinear"}]
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
        # (no map on the analytic sphere: it is one-sided -- face_normal is not flipped towards rays that hit it from inside, and
        # ensure_valid_reflection, core/sampling.art:120-165, then normalises a zero vector: NaN in the reference as well)
        shape = {"d_near": "Box", "d_lin": "Ball", "c_cub": "Ball", "g_tex": "Box", "b_rc": "UV", "n_d": "Box", "n_rc": "UV"}[b]
        s["entities"].append({"name": f"e{k}", "shape": shape, "bsdf": b, "transform": [{"translate": [-2.4 + 0.8 * k, 0.6 * ((k % 2) * 2 - 1), 0]}, {"rotate": [10 * k, 20, 5 * k]}]})
    s["camera"]["transform"] = {"lookat": {"origin": [0.5, -6.5, 3.0], "target": [0, 0, 0], "up": [0, 0, 1]}}
    s["lights"] = [{"type": "env", "name": "env", "radiance": [0.6, 0.7, 0.8]}, {"type": "point", "name": "p", "position": [1, -2, 3], "intensity": [15, 14, 13]}]
    t = _check(s, 240, 160, 4, 2)
    assert len(t.textures) == 6 and len(t.images) == 3
    assert sorted(int(x) for x in t.materials["map_kind"]) == [0, 0, 0, 0, 0, 1, 2, 2]


def test_textured_environment_lights_match_oracle(tmp_path):
    """make_environment_light_textured / make_environment_light over image and checkerboard textures next to finite lights, through all
    three light selectors (the env light takes half of the selector's samples); glossy and diffuse receivers."""
    rng = np.random.default_rng(11)
    env = (rng.random((24, 48, 3)) ** 4 * 5).astype(np.float32)
    np.save(tmp_path / "env.npy", env)
    for sel, cdf, filt in (("uniform", "conditional", "bilinear"), ("simple", "conditional", "bicubic"), ("hierarchy", "none", "nearest"), ("hierarchy", "conditional", "nearest")):
        s = flat_scene()
        s["technique"] = {"type": "path", "max_depth": 4, "light_selector": sel}
        s["film"]["size"] = [160, 120]
        s["textures"] = [{"type": "image", "name": "env", "filename": str(tmp_path / "env.npy"), "filter_type": filt},
                         {"type": "checkerboard", "name": "check", "scale_x": 8, "scale_y": 4, "color0": [0.1, 0.1, 0.3], "color1": [1.5, 1.2, 0.8]}]
        s["bsdfs"] = [{"type": "conductor", "name": "ground", "roughness": 0.4, "material": "silver"}]
        s["lights"] = [{"type": "env", "name": "e", "radiance": "env", "cdf": cdf, "scale": [1.0, 0.5, 2.0]},
                       {"type": "point", "name": "p", "position": [-0.5, 0.3, -1], "intensity": [0.2, 0.2, 0.2]},
                       {"type": "point", "name": "q", "position": [0.5, -0.3, -1.5], "intensity": [0.1, 0.3, 0.2]}]
        _check(s, 160, 120, 2, 2)
    s["lights"][0] = {"type": "env", "name": "e", "radiance": "check"}      # a texture without an image: baked 1 x 1 -> no cdf
    t = _check(s, 160, 120, 2, 2)
    assert int(t.infinite_lights[0]["type"]) == 9


def test_spi1_is_bitwise_reproducible_and_equal_to_oracle():
    s = flat_scene()
    s["lights"].append({"type": "point", "name": "_light", "position": [0, 0, -2], "intensity": [1, 1, 1]})
    t = load_scene(s)
    a, ref, _, _ = render_both(t, 200, 200, 1, 1, seed=42)
    b, _, _, _ = render_both(t, 200, 200, 1, 1, seed=42)
    np.testing.assert_array_equal(a, b)                       # src/tests/integrator/test_reproducibility.py:5-11
    np.testing.assert_array_equal(a, ref)                     # one splat chain per pixel: no reordering possible
    c, _, _, _ = render_both(t, 200, 200, 4, 1, seed=42)
    assert not np.allclose(a, c)                              # test_reproducibility.py:14-20


def test_analytic_scene_averages_on_gpu():
    # src/tests/integrator/test_lights.py:5-44, same configuration (1000^2, 8 iterations, default GPU spi)
    def avg(scene):
        with Runtime(scene) as rt:
            for _ in range(8):
                rt.step()
            return float(rt.image().mean())
    assert avg(flat_scene()) == pytest.approx(0, abs=1e-8)
    s = flat_scene()
    s["lights"].append({"type": "point", "name": "_light", "position": [0, 0, -2], "power": 1})
    assert avg(s) == pytest.approx(0.005100456, abs=1e-4)
    s = flat_scene()
    s["lights"].append({"type": "spot", "name": "_light", "cutoff": 45, "falloff": 45, "position": [0, 0, -2], "direction": [0, 0, 1], "power": 1})
    assert avg(s) == pytest.approx(0.0348280902, abs=2.5e-3)
    s = flat_scene()
    s["lights"].append({"type": "env", "name": "_light", "radiance": [1, 1, 1]})
    assert avg(s) == pytest.approx(1, rel=1e-4)
    assert avg({}) == 0.0                                     # test_init.py:9-12


def test_list_emitter_matches_oracle():
    # igtrace path: Runtime::trace, rays in -> radiance out
    t = load_scene(scene_path("diamond_scene.json"))
    rays = camera_rays(t, 64, 36, seed=9)
    o = Oracle(t)
    ref = o.render(len(rays), 1, spi=1, iteration=0, rays=rays).reshape(-1, 3)
    with Runtime(t, 64, 36, spi=1) as rt:
        got = rt.trace(rays)
    assert rel_l2(got, ref) <= REL_L2_TOL


def test_tile_partition_sums_to_full_frame():
    t = load_scene(scene_path("diamond_scene.json"))
    w, h, spi = 200, 120, 2
    with Runtime(t, w, h, spi=spi) as rt:
        rt.step()
        full = rt.getFramebufferForHost().copy()
    acc = np.zeros_like(full)
    for r in range(3):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setPartition(r, 3, 32)
            rt.step()
            part = rt.getFramebufferForHost().copy()
        assert ((part != 0) & (acc != 0)).sum() == 0      # disjoint support
        acc += part
    assert rel_l2(acc, full) <= 1e-6


@pytest.mark.parametrize("scene,world", [("diamond_scene.json", 1), ("primitives.json", 1), ("synthetic_room.json", 4)])
def test_full_size_configs_match_oracle(scene, world):
    """BASELINE configs C2 / C3 / C4 (stand-in) at THEIR OWN size -- 1920x1080, spi 4, one iteration -- against the oracle:
    relative L2 <= 1e-4 (north_star) and exactly the oracle's camera / shadow / bounce ray counts. C4 is specified on 4 GPUs: its
    frame is rendered as the four tile partitions `setPartition(r, 4)` of a 4-rank run (one after the other on this GPU; the
    partitions are independent, tests/test_multi_gpu.py runs them on separate devices) and summed as the exchange does.
    diamond_scene has no reference image: parity on C2 rests on GPU == oracle plus the oracle's physics pins."""
    w, h, spi = 1920, 1080, 4
    t = load_scene(scene_path(scene))
    o = Oracle(t)
    ref = o.render(w, h, spi=spi, iteration=0)
    acc = np.zeros_like(ref)
    counts = np.zeros(3, np.int64)
    for r in range(world):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setPartition(r, world, 32)
            rt.step()
            part = rt.getFramebufferForHost().copy()
            st = rt.device.getStatistics()
            if world == 1:   # same seed, same iteration -> same image up to the order of the float atomics
                rt.reset()
                rt.step()
                assert rel_l2(rt.getFramebufferForHost(), part) <= 1e-6
        assert ((part != 0) & (acc != 0)).sum() == 0          # the ranks' tiles are disjoint
        acc += part
        counts += (st["CameraRayCount"], st["ShadowRayCount"], st["BounceRayCount"])
    assert np.isfinite(acc).all() and (acc >= 0).all()
    assert counts[0] == w * h * spi
    assert tuple(int(x) for x in counts) == tuple(int(x) for x in o.counters)
    assert rel_l2(acc, ref) <= REL_L2_TOL


def test_deferred_tail_is_invisible():
    """render() may leave the deepest paths of an iteration to the next launch (igb200.h): every observation must still
    see the result of a synchronous render -- same image, exactly the same ray counters, also across a clear."""
    t = load_scene(scene_path("diamond_scene.json"))
    w, h, spi = 320, 180, 4
    out = {}
    for permille, split, fuse in ((0, 0, 1), (50, 2, 1), (500, 0, 3), (4000, 3, 2)):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setOption("defer_permille", permille)
            rt.device.setOption("split_turns", split)   # leading turns as separate shade / trace launches
            rt.device.setOption("fuse", fuse)           # consecutive iterations generated by one launch
            for _ in range(3):
                rt.step()
            img = rt.getFramebufferForHost().copy()
            st = rt.device.getStatistics()
            rt.step()                       # leaves paths in flight ...
            rt.device.clearAllFramebuffer()  # ... which belong to the image being thrown away
            assert not rt.getFramebufferForHost().any()
            rt.IterationCount = 0
            rt.step()
            again = rt.getFramebufferForHost().copy()
        out[permille] = (img, st, again)
    ref_img, ref_st, ref_again = out[0]
    assert ref_st["KernelLaunches"] == 3
    for permille in (50, 500, 4000):
        img, st, again = out[permille]
        assert rel_l2(img, ref_img) <= 1e-6
        assert rel_l2(again, ref_again) <= 1e-6
        for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats"):
            assert st[k] == ref_st[k], k
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(3):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    assert rel_l2(out[4000][0], ref) <= REL_L2_TOL


@pytest.mark.parametrize("scene,w,h,spi,iters,opts", [
    ("diamond_scene.json", 320, 180, 4, 24, {}),                                  # max_depth 64: a frame is finished ~13 launches after its render
    ("diamond_scene.json", 200, 120, 2, 20, {"split_turns": 0, "defer_permille": 50}),
    ("diamond_scene.json", 200, 120, 2, 20, {"fuse": 4}),                         # several iterations per launch (a rank's share at N GPUs)
    ("many_point_lights.json", 160, 160, 1, 10, {}),
    ("evaluation/cbox-d6.json", 128, 128, 2, 12, {"fuse": 1}),                     # max_depth 6: finished after two launches
])
def test_streamed_frames_are_the_synchronous_frames(scene, w, h, spi, iters, opts):
    """igb200_frame_stream_*: frame k handed out while later iterations render == the oracle's sum of iterations 0..k == what the
    synchronous igb200_framebuffer returns after render(k); frames arrive in order, each exactly once, none before it is complete."""
    t = load_scene(scene_path(scene))
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    refs = []
    for it in range(iters):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
        refs.append(ref.copy())
    got = {}
    with Runtime(t, w, h, spi=spi) as rt:
        for k, v in opts.items():
            rt.device.setOption(k, v)
        rt.device.frameStreamBegin()
        early = 0
        for it in range(iters):
            rt.step()
            while (f := rt.device.frameStreamNext(0)) is not None:
                got[f[0]] = f[1].copy()
                early += 1
        while (f := rt.device.frameStreamNext(2)) is not None:
            got[f[0]] = f[1].copy()
        final = rt.device.getFramebufferForHost().copy()          # the synchronous call still works and sees everything
        st = rt.device.getStatistics()
        rt.device.frameStreamEnd()
        rt.step()                                                  # ... and the device renders normally afterwards
        after = rt.device.getFramebufferForHost().copy()
    assert sorted(got) == list(range(iters))
    for it in range(iters):
        assert rel_l2(got[it], refs[it]) <= REL_L2_TOL, it
    assert rel_l2(final, refs[-1]) <= REL_L2_TOL
    assert (st["CameraRayCount"], st["ShadowRayCount"], st["BounceRayCount"]) == tuple(int(x) for x in o.counters)
    o.render(w, h, spi=spi, iteration=iters, fb=ref)
    assert rel_l2(after, ref) <= REL_L2_TOL


@pytest.mark.parametrize("scene,w,h,spi", [("many_point_lights.json", 640, 640, 1), ("diamond_scene.json", 480, 270, 4)])
def test_material_binning_is_invisible(scene, w, h, spi):
    """a9: shading through the per-class index lists the trace phase writes (split turns) gives the image and the exact ray counts of
    shading in queue order -- only the order of the float additions into a pixel can differ. Also with a deferred tail and fused iterations."""
    t = load_scene(scene_path(scene))
    out = {}
    for mode, extra in ((0, {}), (1, {}), (1, {"split_turns": 6, "defer_permille": 300}), (1, {"fuse": 2})):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setOption("bin_materials", mode)
            for k, v in extra.items():
                rt.device.setOption(k, v)
            for _ in range(3):
                rt.step()
            img = rt.getFramebufferForHost().copy()
            st = rt.device.getStatistics()
        out[(mode, tuple(extra))] = (img, st)
    ref_img, ref_st = out[(0, ())]
    for key, (img, st) in out.items():
        assert rel_l2(img, ref_img) <= 1e-6, key
        for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats"):
            assert st[k] == ref_st[k], (key, k)
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(3):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    assert rel_l2(out[(1, ())][0], ref) <= REL_L2_TOL


@pytest.mark.parametrize("scene", ["diamond_scene.json", "evaluation/cbox-d6.json", "many_point_lights.json", "evaluation/multilight-uniform.json"])
def test_merged_tree_walk_is_invisible(scene):
    """Small scenes are traced through one merged tree (instances refitted in world space, traverse.cuh): same image, exactly the same
    ray counts as the two-level walk, and both equal to the oracle."""
    t = load_scene(scene_path(scene))
    w, h, spi = 320, 200, 2
    out = {}
    for flat in (1, 0):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setOption("flat", flat)
            for _ in range(3):
                rt.step()
            out[flat] = (rt.getFramebufferForHost().copy(), rt.device.getStatistics())
    assert rel_l2(out[1][0], out[0][0]) <= 1e-6
    for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats"):
        assert out[1][1][k] == out[0][1][k], k
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(3):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    assert rel_l2(out[1][0], ref) <= REL_L2_TOL
    assert (out[1][1]["CameraRayCount"], out[1][1]["ShadowRayCount"], out[1][1]["BounceRayCount"]) == tuple(int(x) for x in o.counters)


def test_white_furnace_through_glass_on_gpu():
    """Energy conservation of the whole device pipeline on delta paths (tests/test_oracle_kat.py has the oracle's)."""
    t = load_scene(furnace_scene())
    with Runtime(t, 256, 256, spi=8) as rt:
        for _ in range(8):
            rt.step()
        img = rt.image()
        st = rt.device.getStatistics()
    assert st["BounceRayCount"] > 0.2 * st["CameraRayCount"]
    assert float(img.mean()) == pytest.approx(1.0, abs=2e-3)
    assert np.abs(img.reshape(16, 16, 16, 16, 3).mean(axis=(1, 3, 4)) - 1).max() < 0.05


@pytest.mark.parametrize("w,h,spi,seed,world", [(97, 53, 3, 7, 1), (33, 129, 5, 123, 3), (1, 1, 1, 0, 1), (300, 7, 2, 99, 2)])
def test_odd_sizes_seeds_and_partitions(w, h, spi, seed, world):
    """Ragged frames (not multiples of the 32-pixel tile), odd spi, user seeds, ranks with unequal tile counts, fused iterations."""
    t = load_scene(scene_path("diamond_scene.json"))
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(3):
        o.render(w, h, spi=spi, iteration=it, seed=seed, fb=ref)
    acc = np.zeros_like(ref)
    cam = 0
    for r in range(world):
        with Runtime(t, w, h, spi=spi, seed=seed) as rt:
            rt.device.setPartition(r, world, 32)
            rt.device.setOption("fuse", 2)
            for _ in range(3):
                rt.step()
            acc += rt.getFramebufferForHost()
            cam += rt.device.getStatistics()["CameraRayCount"]
    assert cam == w * h * spi * 3
    assert rel_l2(acc, ref) <= REL_L2_TOL


def test_errors_are_reported_not_rendered():
    t = load_scene(scene_path("single_triangle.json"))
    with B200Device() as dev:
        with pytest.raises(Exception, match="no scene"):
            dev.render(1, 8, 8, 0)
        dev.assignScene(t)
        with pytest.raises(Exception, match="overflows"):
            dev.render(1 << 20, 4096, 4096, 0)
        with pytest.raises(Exception, match="spi"):
            dev.render(0, 8, 8, 0)
        with pytest.raises(Exception, match="AOV"):
            dev.getFramebufferForHost("Normals")
        bad = load_scene(scene_path("single_triangle.json"))
        bad.materials["bsdf"][0] = 7
        with pytest.raises(Exception, match="unsupported bsdf"):
            dev.assignScene(bad)
        bad = load_scene(scene_path("single_triangle.json"))
        bad.materials["tex"][0, 0] = 3
        with pytest.raises(Exception, match="texture that does not exist"):
            dev.assignScene(bad)
        bad = load_scene(scene_path("many_point_lights.json"))
        bad.aux_data = bad.aux_data[:1000]
        with pytest.raises(Exception, match="does not fit aux_data"):
            dev.assignScene(bad)
        bad = load_scene(scene_path("single_triangle.json"))     # a trimesh header that claims more faces than the table holds
        bad.shape_data = bad.shape_data.copy()
        bad.shape_data[:4] = np.frombuffer(np.uint32(1 << 20).tobytes(), np.uint8)
        with pytest.raises(Exception, match="does not fit the shapes table"):
            dev.assignScene(bad)
        dev.render(1, 8, 8, 0)   # the device is still usable after the errors
        assert np.isfinite(dev.getFramebufferForHost()).all()


def test_gold_conductor_matches_oracle():
    scene = furnace_scene()
    scene["bsdfs"] = [{"type": "conductor", "name": "glass", "material": "gold"}]
    scene["lights"].append({"type": "point", "name": "p", "position": [3, -2, 4], "intensity": [20, 20, 20]})
    t = load_scene(scene)
    got, ref, stats, cnt = render_both(t, 128, 128, 4, 2)
    assert rel_l2(got, ref) <= REL_L2_TOL
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in cnt)


def test_spot_lights_match_oracle():
    s = flat_scene()
    s["lights"].append({"type": "spot", "name": "a", "cutoff": 45, "falloff": 30, "position": [0, 0, -2], "direction": [0.1, 0, 1], "power": [1, 2, 3]})
    s["lights"].append({"type": "spot", "name": "b", "cutoff": 20, "falloff": 20, "position": [0.5, 0, -2], "direction": [0, 0, 1], "intensity": [1, 1, 1]})
    s["lights"].append({"type": "point", "name": "c", "position": [-0.5, 0.3, -1], "intensity": [0.2, 0.2, 0.2]})
    t = load_scene(s)
    got, ref, stats, cnt = render_both(t, 200, 200, 2, 2)
    assert rel_l2(got, ref) <= REL_L2_TOL
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in cnt)


def test_distant_lights_match_oracle():
    """Directional (delta) and sun (cone) lights next to an environment, all three light selectors."""
    for sel in ("uniform", "simple", "hierarchy"):
        s = flat_scene()
        s["technique"]["light_selector"] = sel
        s["lights"].append({"type": "directional", "name": "d", "direction": [0.2, 0.1, 1], "irradiance": [1, 0.5, 0.25]})
        s["lights"].append({"type": "sun", "name": "s", "direction": [0.3, 0.2, -1], "irradiance": [3, 2, 1], "angle": 2.5})
        s["lights"].append({"type": "env", "name": "e", "radiance": [0.1, 0.1, 0.1]})
        s["lights"].append({"type": "point", "name": "p", "position": [-0.5, 0.3, -1], "intensity": [0.2, 0.2, 0.2]})
        s["lights"].append({"type": "point", "name": "q", "position": [0.5, -0.3, -1.5], "intensity": [0.1, 0.3, 0.2]})
        t = load_scene(s)
        got, ref, stats, cnt = render_both(t, 200, 200, 2, 2)
        assert ref.sum() > 0
        assert rel_l2(got, ref) <= REL_L2_TOL, sel
        assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in cnt), sel


@pytest.mark.parametrize("selector", ["simple", "hierarchy"])
def test_many_lights_selectors_match_oracle(selector):
    """25 finite lights + an environment: a light tree five levels deep, codes of the pdf walk, the 50 % infinite-light branch."""
    from test_light_selectors import _many_lights_scene
    t = load_scene(_many_lights_scene(selector))
    got, ref, stats, cnt = render_both(t, 160, 120, 2, 2)
    assert ref.sum() > 0
    assert rel_l2(got, ref) <= REL_L2_TOL
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in cnt)


def test_standard_aovs_match_oracle():
    """Normals / Albedo AOVs of the reference's infobuffer wrapper (technique/internal/infobuffer.art): first-hit shading normal and
    BSDF albedo, written at iteration 0 only, through the C ABI and through the C++ plugin (which finds the wrapper in the script)."""
    from ignis_b200 import plugin
    t = load_scene(scene_path("diamond_scene.json"))
    w, h, spi = 160, 90, 4
    o = Oracle(t)
    ref_n, ref_a, ref = (np.zeros((h, w, 3), np.float32) for _ in range(3))
    o.set_aovs(ref_n, ref_a)
    for it in range(2):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    assert np.abs(ref_n).sum() > 0 and ref_a.max() <= 1.0 + 1e-6
    with Runtime(t, w, h, spi=spi) as rt:
        with pytest.raises(Exception, match="AOV"):
            rt.device.getFramebufferForHost("Normals")          # off by default
        rt.device.setOption("std_aovs", 1)
        rt.step(); rt.step()
        got = [rt.device.getFramebufferForHost(n).copy() for n in ("Normals", "Albedo", "")]
    with plugin.PluginRuntime(t, w, h, spi, std_aovs=True) as prt:
        prt.step(); prt.step()
        got_p = [prt.getFramebufferForHost(n).copy() for n in ("Normals", "Albedo", "")]
    for g in (got, got_p):
        assert rel_l2(g[0], ref_n) <= 1e-5 and rel_l2(g[1], ref_a) <= 1e-5 and rel_l2(g[2], ref) <= REL_L2_TOL
    with plugin.PluginRuntime(t, w, h, spi, std_aovs=False) as prt:
        prt.step()
        with pytest.raises(Exception, match="AOV"):
            prt.getFramebufferForHost("Albedo")


@pytest.mark.gpu
@pytest.mark.parametrize("scene,w,h,spi,iters,world", [
    ("diamond_scene.json", 240, 135, 4, 3, 1),               # C2 at reduced size, the reference GPU default spi
    ("primitives.json", 200, 120, 3, 2, 1),
    ("many_point_lights.json", 160, 160, 2, 2, 1),           # textures, rough conductor, sky: the full shade kernels
    ("evaluation/cbox-d6.json", 128, 128, 8, 2, 3),           # three partitions summed
])
def test_deterministic_accumulation_is_bit_exact(scene, w, h, spi, iters, world):
    """Option "deterministic": every sample sums its own contributions in path order, the samples are folded into the pixel in sample
    order. Same inputs => the SAME BITS, run to run (src/tests/integrator/test_reproducibility.py:5-11 at any spi) -- and the same bits as
    the oracle accumulating the same way, so the 1e-7 that float atomics leave in the other tests is gone."""
    t = load_scene(scene_path(scene))
    o = Oracle(t)
    o.set_deterministic(True)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(iters):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    frames = []
    for run in range(2):
        acc = np.zeros((h, w, 3), np.float32)
        for rank in range(world):
            with Runtime(t, w, h, spi=spi) as rt:
                rt.device.setOption("deterministic", 1)
                if world > 1:
                    rt.device.setPartition(rank, world, 32)
                for _ in range(iters):
                    rt.step()
                acc += rt.getFramebufferForHost()            # disjoint supports: x + 0 is exact
        frames.append(acc)
    np.testing.assert_array_equal(frames[0].view(np.uint32), frames[1].view(np.uint32))
    np.testing.assert_array_equal(frames[0].view(np.uint32), ref.view(np.uint32))
    # and it is the same image as the default accumulation up to the reordering of float additions
    with Runtime(t, w, h, spi=spi) as rt:
        for _ in range(iters):
            rt.step()
        plain = rt.getFramebufferForHost().copy()
    assert rel_l2(plain, ref) <= 1e-5
    if spi > 1 and scene == "diamond_scene.json":
        assert not np.array_equal(plain.view(np.uint32), ref.view(np.uint32))   # ... which is what the option is for


@pytest.mark.gpu
def test_deterministic_aovs_and_exclusions():
    t = load_scene(scene_path("diamond_scene.json"))
    w, h, spi = 160, 90, 4
    o = Oracle(t)
    o.set_deterministic(True)
    ref_n, ref_a, ref = (np.zeros((h, w, 3), np.float32) for _ in range(3))
    o.set_aovs(ref_n, ref_a)
    for it in range(2):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    with Runtime(t, w, h, spi=spi) as rt:
        rt.device.setOption("std_aovs", 1)
        rt.device.setOption("deterministic", 1)
        rt.step(); rt.step()
        got = [rt.device.getFramebufferForHost(n).copy() for n in ("Normals", "Albedo", "")]
        with pytest.raises(Exception, match="exclude"):
            rt.device.frameStreamBegin(4)
    for g, r in zip(got, (ref_n, ref_a, ref)):
        np.testing.assert_array_equal(g.view(np.uint32), r.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["diamond_scene.json", "evaluation/cbox-d6.json", "many_point_lights.json"])
def test_staged_ray_records_are_invisible(scene):
    """The merged-tree trace kernel fed through shared memory (TMA bulk copies of the ray records one batch ahead, wavefront.cuh
    phase_trace_staged; CTAs of 384 or 768 threads) against the one that reads the records from global memory: same hits, hence
    bit-identical frames in deterministic mode and identical ray counts."""
    t = load_scene(scene_path(scene))
    w, h, spi = 333, 187, 3     # odd sizes: the last batch is partial and one batch straddles the primary / shadow boundary
    out = {}
    for blk in (256, 384, 768):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setOption("flat_block", blk)
            rt.device.setOption("deterministic", 1)
            for _ in range(3):
                rt.step()
            out[blk] = (rt.getFramebufferForHost().copy(), rt.device.getStatistics())
    for blk in (384, 768):
        np.testing.assert_array_equal(out[blk][0].view(np.uint32), out[256][0].view(np.uint32))
        for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats"):
            assert out[blk][1][k] == out[256][1][k], (blk, k)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,w,h", [("diamond_scene.json", 240, 135), ("primitives.json", 240, 135), ("synthetic_room.json", 192, 108), ("evaluation/room.json", 128, 128)])
def test_gpu_built_bvh_is_invisible(scene, w, h):
    """the shapes' BVH8s built ON THE GPU (Morton codes, radix tree, bottom-up boxes, collapse; csrc/bvh_build.cu) instead of by the
    host's binned-SAH builder. The closest hit is a pure function of the ray, so hit records are the same bits and deterministic renders are
    bit-identical, whatever the tree looks like -- and both equal the oracle's brute-force answer."""
    t = load_scene(scene_path(scene))
    rays = camera_rays(t, 96, 54)
    out = {}
    for gpu in (0, 1):
        with Runtime(t, w, h, spi=2) as rt:
            rt.device.setOption("gpu_bvh", gpu)
            rt.device.setOption("deterministic", 1)
            rt.device.assignScene(t)
            info = rt.device.sceneBuildInfo()
            for _ in range(2):
                rt.step()
            out[gpu] = (rt.getFramebufferForHost().copy(), rt.device.traceClosest(rays), rt.device.traceAny(rays), rt.device.getStatistics(), info)
    assert out[0][4]["gpu_built"] == 0 and out[1][4]["gpu_built"] >= 1 and out[1][4]["nodes"] > 0
    np.testing.assert_array_equal(out[1][0].view(np.uint32), out[0][0].view(np.uint32))
    for k in ("ent_id", "prim_id"):
        np.testing.assert_array_equal(out[1][1][k], out[0][1][k])
    for k in ("t", "u", "v"):
        np.testing.assert_array_equal(out[1][1][k].view(np.uint32), out[0][1][k].view(np.uint32))
    np.testing.assert_array_equal(out[1][2], out[0][2])
    for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats"):
        assert out[1][3][k] == out[0][3][k], k
    ref = Oracle(t).trace_closest(rays, use_bvh=False)    # brute force
    np.testing.assert_array_equal(out[1][1]["prim_id"], ref["prim_id"])
    np.testing.assert_array_equal(out[1][1]["t"].view(np.uint32), ref["t"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["synthetic_room.json", "primitives.json"])
def test_merged_tree_in_global_memory_is_invisible(scene):
    """Option "flat" = 2: scenes too large for shared memory are traced through the merged single-level tree as well (every instance's
    nodes refitted in world space, read through L1 / L2): bit-identical frames and ray counts against the two-level walk."""
    t = load_scene(scene_path(scene))
    w, h, spi = 320, 180, 2
    out = {}
    for flat in (2, 0):
        with Runtime(t, w, h, spi=spi) as rt:
            rt.device.setOption("flat", flat)
            rt.device.setOption("deterministic", 1)
            rt.device.assignScene(t)
            for _ in range(2):
                rt.step()
            out[flat] = (rt.getFramebufferForHost().copy(), rt.device.getStatistics())
    np.testing.assert_array_equal(out[2][0].view(np.uint32), out[0][0].view(np.uint32))
    for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats"):
        assert out[2][1][k] == out[0][1][k], k


@pytest.mark.gpu
@pytest.mark.parametrize("gpu", [0, 1])
def test_bvh_cache_round_trip(tmp_path, gpu):
    """The on-disk BVH cache (the reference: TriMeshProvider.cpp:326-351): first load builds and stores, second load reads the files,
    a corrupted file is ignored and rebuilt; the image never changes."""
    t = load_scene(scene_path("synthetic_room.json"))
    w, h = 160, 90
    def run(expect):
        with Runtime(t, w, h, spi=1) as rt:
            rt.device.setOption("gpu_bvh", gpu)
            rt.device.setOption("bvh_cache_min_faces", 1000)     # the room's icospheres have 20 480 / 81 920 faces
            rt.device.setOption("deterministic", 1)
            rt.device.setCacheDir(tmp_path)
            rt.device.assignScene(t)
            info = rt.device.sceneBuildInfo()
            rt.step()
            img = rt.getFramebufferForHost().copy()
        for k, v in expect.items():
            assert info[k] == v, (k, info)
        return img, info
    a, info = run({"cache_loaded": 0})
    n = info["cache_stored"]
    assert n >= 2 and len(list(tmp_path.glob("bvh8_*.bin"))) == n
    b, _ = run({"cache_loaded": n, "cache_stored": 0, "gpu_built": 0, "host_built": 2})   # the two rectangles (<= 4 faces) are never cached
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    victim = sorted(tmp_path.glob("bvh8_*.bin"))[0]
    raw = bytearray(victim.read_bytes()); raw[200] ^= 0xFF; raw[-3] ^= 0xFF            # a child code and an order entry
    victim.write_bytes(bytes(raw[: len(raw) - 8]))                                       # ... and truncated
    c, info = run({"cache_loaded": n - 1, "cache_stored": 1})
    np.testing.assert_array_equal(a.view(np.uint32), c.view(np.uint32))
    with Runtime(t, w, h, spi=1) as rt:                                                  # no cache directory: nothing is read or written
        rt.device.setOption("bvh_cache_min_faces", 1000)
        rt.device.assignScene(t)
        assert rt.device.sceneBuildInfo()["cache_loaded"] == 0 and rt.device.sceneBuildInfo()["cache_stored"] == 0
