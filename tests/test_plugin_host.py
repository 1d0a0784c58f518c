"""The C++ plugin layer (csrc/host): the compiler device recognises the reference's stage scripts, the render device binds
them to the C ABI. CPU part: script text -> descriptors must reproduce, bit for bit, the descriptors the scene loader
computed directly. GPU part (`-m gpu`): rendering through ig_get_interface() / IRenderDevice equals the oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest

from ignis_b200 import plugin, refscript
from ignis_b200.scene import load_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = ["single_triangle.json", "diamond_scene.json", "primitives.json", "evaluation/cbox-d6.json", "evaluation/multilight-uniform.json",
          "evaluation/multilight-simple.json", "evaluation/multilight-hierarchy.json",
          "evaluation/emissive-plane.json", "evaluation/point.json", "evaluation/plane-d1.json", "evaluation/sphere-light-pure.json",
          "evaluation/two-planes-mirror.json", "evaluation/sun-on-plane.json", "<spot>", "<distant>", "<procedural>", "<points-only>", "<bitmaps>",
          "many_point_lights.json", "evaluation/env4k-conditional.json", "evaluation/env4k-none.json"]


def _write_png(path, px):
    """(H, W, C) u8 -> an 8-bit PNG whose scanlines use filters 0..4 in rotation (so a reader is tested on every one of them)."""
    import struct, zlib
    h, w, ch = px.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    raw = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        cur = px[y].reshape(-1).astype(np.int32)
        ft = y % 5
        a = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        c = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
        if ft == 0: line = cur
        elif ft == 1: line = cur - a
        elif ft == 2: line = cur - prev
        elif ft == 3: line = cur - ((a + prev) >> 1)
        else:
            pa, pb, pc = np.abs(prev - c), np.abs(a - c), np.abs(a + prev - 2 * c)
            line = cur - np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        raw += bytes([ft]) + (line & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    comp = zlib.compress(bytes(raw), 9)
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + chunk(b"IDAT", comp[:len(comp) // 2])
                 + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b""))


def scene(name):
    if name == "<spot>":   # the spot-light scene of the reference's integrator tests (src/tests/integrator/test_lights.py:25-37)
        from conftest import flat_scene
        s = flat_scene()
        s["lights"].append({"type": "spot", "name": "_light", "cutoff": 45, "falloff": 30, "position": [0, 0, -2], "direction": [0.1, 0, 1], "power": [1, 2, 3]})
        s["lights"].append({"type": "spot", "name": "_light2", "cutoff": 20, "falloff": 20, "position": [0.5, 0, -2], "direction": [0, 0, 1], "intensity": [1, 1, 1]})
        return load_scene(s)
    if name == "<distant>":   # a directional light and a sun given by irradiance and elevation / azimuth next to an environment
        from conftest import flat_scene
        s = flat_scene()
        s["lights"].append({"type": "directional", "name": "_d", "direction": [0.2, 0.1, 1], "irradiance": [1, 0.5, 0.25]})
        s["lights"].append({"type": "sun", "name": "_s", "elevation": 1.1, "azimuth": 0.4, "irradiance": [3, 2, 1], "angle": 2.5})
        s["lights"].append({"type": "env", "name": "_e", "radiance": [0.1, 0.1, 0.1]})
        return load_scene(s)
    if name == "<procedural>":
        # everything of SURVEY 8f-1 / f-2 that needs no image file: checkerboard textures (one with a transform) as colours and as bump / normal
        # maps, rough conductors (explicit and roughness + anisotropic), and 12 simple point lights -> an embedded fix-table
        from conftest import furnace_scene
        s = furnace_scene()
        s["technique"]["max_depth"] = 6
        s["textures"] = [{"type": "checkerboard", "name": "check", "scale_x": 6, "scale_y": 3, "color0": [0.2, 0.3, 0.4], "color1": [0.9, 0.8, 0.7], "transform": {"rotate": [0, 0, 30]}},
                         {"type": "checkerboard", "name": "grid", "scale_x": 11, "scale_y": 11, "color0": [0, 0, 0], "color1": [1, 1, 1]}]
        s["bsdfs"] = [{"type": "diffuse", "name": "d_check", "reflectance": "check"},
                      {"type": "conductor", "name": "rc", "roughness": 0.2, "anisotropic": 0.4, "material": "copper"},
                      {"type": "conductor", "name": "rc2", "roughness_u": 0.15, "roughness_v": 0.3, "specular_reflectance": "grid"},
                      {"type": "dielectric", "name": "g_tex", "specular_reflectance": "check", "specular_transmittance": "grid"},
                      {"type": "bumpmap", "name": "b_rc", "bsdf": "rc", "map": "grid", "strength": 0.35},
                      {"type": "normalmap", "name": "n_d", "bsdf": "d_check", "map": "check", "strength": 0.8},
                      {"type": "diffuse", "name": "white", "reflectance": [1, 1, 1]}]
        s["shapes"] = [{"type": "cube", "name": "Box", "width": 1.0, "height": 1.0, "depth": 1.0, "origin": [-0.5, -0.5, -0.5]},
                       {"type": "rectangle", "name": "Floor", "width": 12, "height": 12, "origin": [-6, -6, -0.9]},
                       {"type": "rectangle", "name": "Lamp", "width": 1, "height": 1, "flip_normals": True}, {"type": "uvsphere", "name": "UV", "radius": 0.45}]
        s["entities"] = [{"name": "Floor", "shape": "Floor", "bsdf": "d_check"},
                         {"name": "Lamp", "shape": "Lamp", "bsdf": "white", "transform": {"translate": [0, 0, 3]}}]
        for k, b in enumerate(["rc", "rc2", "g_tex", "b_rc", "n_d"]):
            s["entities"].append({"name": f"e{k}", "shape": "UV" if k % 2 else "Box", "bsdf": b, "transform": [{"translate": [-2.0 + k, 0.3 * k, 0]}, {"rotate": [10 * k, 20, 5 * k]}]})
        s["camera"]["transform"] = {"lookat": {"origin": [0.5, -6.5, 3.0], "target": [0, 0, 0], "up": [0, 0, 1]}}
        s["lights"] = [{"type": "env", "name": "env", "radiance": [0.3, 0.35, 0.4]}]
        s["lights"] += [{"type": "point", "name": f"p{k}", "position": [np.cos(k) * 3, np.sin(k) * 3, 2 + 0.1 * k], "intensity": [1 + k, 2, 3]} for k in range(12)]
        return load_scene(s)
    if name == "<bitmaps>":
        # 8-bit image files through the script path: the reference's own bump map (a grey PNG) under a rough conductor -- the material of BASELINE
        # config C5 -- and colour PNGs written here (RGB, RGBA, grey + alpha; every scanline filter) with each filter / border / linear variant
        import tempfile
        from conftest import furnace_scene
        tmp = tempfile.mkdtemp(prefix="igb200_png_")
        rng = np.random.default_rng(3)
        files = {}
        for tag, ch in (("rgb", 3), ("rgba", 4), ("ga", 2)):
            files[tag] = os.path.join(tmp, tag + ".png")
            _write_png(files[tag], rng.integers(0, 256, (13, 17, ch), dtype=np.uint8))
        s = furnace_scene()
        s["technique"]["max_depth"] = 5
        s["textures"] = [{"type": "bitmap", "name": "bump", "filename": os.path.join(ROOT, "scenes", "textures/bumpmap.png"), "filter_type": "trilinear"},
                         {"type": "image", "name": "rgb", "filename": files["rgb"], "filter_type": "nearest", "wrap_mode": "mirror"},
                         {"type": "image", "name": "rgba", "filename": files["rgba"], "filter_type": "bilinear", "wrap_mode_u": "clamp", "wrap_mode_v": "repeat", "transform": {"scale": [2.5, 1.5, 1]}},
                         {"type": "image", "name": "ga", "filename": files["ga"], "linear": True},
                         {"type": "image", "name": "rgb_lin", "filename": files["rgb"], "linear": True, "filter_type": "bilinear"}]
        s["bsdfs"] = [{"type": "conductor", "name": "rc", "roughness": 0.2, "material": "gold"},
                      {"type": "bumpmap", "name": "b_rc", "bsdf": "rc", "map": "bump", "strength": 0.5},
                      {"type": "diffuse", "name": "d_rgb", "reflectance": "rgb"}, {"type": "diffuse", "name": "d_rgba", "reflectance": "rgba"},
                      {"type": "dielectric", "name": "g", "specular_reflectance": "ga", "specular_transmittance": "rgb_lin"}]
        s["shapes"] = [{"type": "cube", "name": "Box", "width": 1.0, "height": 1.0, "depth": 1.0, "origin": [-0.5, -0.5, -0.5]},
                       {"type": "rectangle", "name": "Floor", "width": 12, "height": 12, "origin": [-6, -6, -0.9]}, {"type": "uvsphere", "name": "UV", "radius": 0.45}]
        s["entities"] = [{"name": "Floor", "shape": "Floor", "bsdf": "b_rc"}]
        for k, b in enumerate(["d_rgb", "d_rgba", "g"]):
            s["entities"].append({"name": f"e{k}", "shape": "UV" if k % 2 else "Box", "bsdf": b, "transform": [{"translate": [-1.2 + 1.2 * k, 0.2 * k, 0]}, {"rotate": [10 * k, 20, 5 * k]}]})
        s["camera"]["transform"] = {"lookat": {"origin": [0.5, -5.5, 2.5], "target": [0, 0, 0], "up": [0, 0, 1]}}
        s["lights"] = [{"type": "env", "name": "env", "radiance": [0.6, 0.7, 0.8]}, {"type": "point", "name": "p", "position": [1, -2, 3], "intensity": [15, 14, 13]}]
        return load_scene(s)
    if name == "<points-only>":   # nothing but embedded simple point lights: `let finite_lights = e_simplepointlight;` (LoaderLight.cpp:188-193)
        from conftest import flat_scene
        s = flat_scene()
        s["lights"] = [{"type": "point", "name": f"p{k}", "position": [0.1 * k - 0.5, 0.05 * k, -2], "power" if k % 2 else "intensity": [1, 1 + k, 1]} for k in range(11)]
        return load_scene(s)
    return load_scene(os.path.join(ROOT, "scenes", name))


def test_host_library_exports_plugin_entry_points():
    L = plugin.lib()
    for s in plugin.SYMBOLS:
        assert hasattr(L, s), s
    major, minor = C.c_int(), C.c_int()
    assert L.igbh_interface_version(C.byref(major), C.byref(minor)) == 0   # architecture = Nvidia (replaces ig_device_cuda)
    assert (major.value, minor.value) == (0, 3)                            # Build::getVersion() check of DeviceManager.cpp:180-194


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("mode", ["default", "disable", "force"])
def test_recognised_descriptors_equal_loader_descriptors(name, mode, tmp_path):
    t = scene(name)
    st = refscript.generate(t, specialization=mode, cache_dir=str(tmp_path))
    g = plugin.Params(st.global_registry)
    exact = mode != "force"   # `force` prints every value with std::to_string's 6 decimals (ShadingTree.cpp:933-981): lossy by design
    hits = [plugin.CompiledStage(s) for s in st.hits]
    textures = plugin.TextureTable(st.resource_map)
    db = plugin.FixTableDB(st.fix_tables)
    miss_lights = plugin.CompiledStage(st.miss).lights(g, db, textures)   # lights first: the textures of environment lights lead the table, as in the loader
    for i, h in enumerate(hits):
        m = h.material_tex(g, textures)
        assert int(m["bsdf"]) == int(t.materials[i]["bsdf"]) and int(m["light_id"]) == int(t.materials[i]["light_id"])
        for f in ("tex", "distribution", "map_kind", "map_tex"):   # textures by index of first use, the microfacet distribution, the map wrapper
            np.testing.assert_array_equal(m[f], t.materials[i][f], err_msg=f"material {i}: {f}")
        for f in ("alpha_u", "alpha_v", "map_strength"):
            if exact:
                assert np.float32(m[f]).view(np.uint32) == np.float32(t.materials[i][f]).view(np.uint32), (i, f, m[f], t.materials[i][f])
            else:
                np.testing.assert_allclose(m[f], t.materials[i][f], atol=1e-6)
        ref_p = t.materials[i]["p"].copy()
        if int(m["bsdf"]) == 2 and mode == "disable":
            ref_p[9] = 0   # eta / k come from the registry, Artic's `?eta` is false: the reference itself takes make_pure_conductor_bsdf then
        if exact:
            np.testing.assert_array_equal(m["p"].view(np.uint32), ref_p.view(np.uint32))
        else:
            np.testing.assert_allclose(m["p"], ref_p, atol=1e-6)
    got_tex = textures.records()
    assert len(got_tex) == len(t.textures)
    for k in range(len(got_tex)):
        for f in ("type", "image", "filter", "border_u", "border_v"):
            assert int(got_tex[k][f]) == int(t.textures[k][f]), (k, f)
        for f in ("transform", "p"):
            if exact:
                np.testing.assert_array_equal(got_tex[k][f].view(np.uint32), t.textures[k][f].view(np.uint32), err_msg=f"texture {k}: {f}")
            else:
                np.testing.assert_allclose(got_tex[k][f], t.textures[k][f], atol=1e-6)
    got_img = textures.images()   # decoded by the host layer's own PNG reader == what the loader's reader produced
    assert len(got_img) == len(t.images)
    for (gf, ga), (rf, ra) in zip(got_img, t.images):
        assert gf == rf and ga.shape == ra.shape
        np.testing.assert_array_equal(ga, ra)
    np.testing.assert_array_equal(textures.aux().view(np.uint32), np.asarray(t.aux_data, np.float32).view(np.uint32))   # the environment cdfs, read back from their files
    miss_stage = plugin.CompiledStage(st.miss)
    for stage in [miss_stage] + hits[:1]:
        inf, fin = miss_lights if stage is miss_stage else stage.lights(g, db, plugin.TextureTable(st.resource_map))   # (every hit stage repeats the light tables)
        assert len(inf) == len(t.infinite_lights) and len(fin) == len(t.finite_lights)
        for got, ref in list(zip(inf, t.infinite_lights)) + list(zip(fin, t.finite_lights)):
            assert int(got["type"]) == int(ref["type"])
            if int(ref["type"]) in (3, 4):
                assert int(got["entity_id"]) == int(ref["entity_id"])
            if exact:
                np.testing.assert_array_equal(got["p"].view(np.uint32), ref["p"].view(np.uint32))
            else:
                np.testing.assert_allclose(got["p"], ref["p"], rtol=1e-5, atol=1e-6)
        tech = stage.technique(g)
        assert tech.tobytes() == t.technique.tobytes()
        # the buffer of the cdf / hierarchy light selector, read from the file the script names
        np.testing.assert_array_equal(stage.selector_data.view(np.uint32), t.selector_data.view(np.uint32))
    cam = plugin.CompiledStage(st.raygen).camera(g)
    assert cam.tobytes() == t.camera.tobytes()


def test_unknown_constructs_fail_loudly():
    t = scene("diamond_scene.json")
    st = refscript.generate(t)
    g = plugin.Params(st.global_registry)
    # a BSDF the device does not implement must not be rendered as something else
    bad = refscript.Stage("ig_hit_shader", st.hits[0].script.replace("make_diffuse_bsdf(ctx.surf,", "make_principled_bsdf(ctx.surf,"), st.hits[0].local)
    with pytest.raises(plugin.DeviceError, match="make_principled_bsdf"):
        plugin.CompiledStage(bad).material(g)
    # a light selector the device does not implement; a selector whose buffer does not exist
    bad = refscript.Stage("ig_miss_shader", st.miss.script.replace("make_uniform_light_selector(infinite_lights, finite_lights)", "make_power_light_selector(infinite_lights, finite_lights)"), st.miss.local)
    with pytest.raises(plugin.DeviceError, match="make_power_light_selector"):
        plugin.CompiledStage(bad).technique(g)
    bad = refscript.Stage("ig_miss_shader", st.miss.script.replace("make_uniform_light_selector(infinite_lights, finite_lights)", 'make_hierarchy_light_selector(infinite_lights, finite_lights, device.load_buffer("/nonexistent/light_hierarchy.bin"))'), st.miss.local)
    with pytest.raises(plugin.DeviceError, match="cannot open"):
        plugin.CompiledStage(bad).technique(g)
    # embedded light tables (>= 10 simple lights) are rejected at compile time
    bad = refscript.Stage("ig_miss_shader", st.miss.script.replace("let finite_lights = LightTable {", "let finite_lights = load_simple_point_lights(12, 0, device); let unused_table = LightTable {"), st.miss.local)
    with pytest.raises(plugin.DeviceError, match="LightTable"):
        plugin.CompiledStage(bad)
    # entry points that do not exist
    with pytest.raises(plugin.DeviceError, match="not found"):
        plugin.CompiledStage(refscript.Stage("ig_hit_shader", "fn other() -> () { }", refscript.Registry()))
    # stages of the pipeline that carry no parameters still yield a non-null handle (Runtime.cpp:631-657)
    trav = refscript.Stage("ig_traversal_shader", "#[export] fn ig_traversal_shader(settings: &Settings, size: i32, is_secondary: bool) -> () {\n  let spi = settings.spi;\n}\n", refscript.Registry())
    assert plugin.CompiledStage(trav).handle


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spi", [("diamond_scene.json", 160, 90, 2), ("primitives.json", 160, 90, 2), ("evaluation/multilight-uniform.json", 96, 96, 2),
                                           ("evaluation/multilight-hierarchy.json", 96, 96, 2), ("evaluation/multilight-simple.json", 96, 96, 2),
                                           ("<procedural>", 160, 120, 2), ("<points-only>", 96, 96, 2), ("<bitmaps>", 160, 120, 2)])
def test_render_through_cpp_plugin_matches_oracle(name, w, h, spi):
    from oracle.oracle import Oracle
    t = scene(name)
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(2):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    with plugin.PluginRuntime(t, w, h, spi) as rt:
        rt.step()
        rt.step()
        got = rt.getFramebufferForHost().copy()
        stats = rt.getStatistics()
    err = float(np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel()))
    assert err <= 1e-4, err
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in o.counters)


@pytest.mark.gpu
def test_trace_through_cpp_plugin_matches_direct_abi():
    from ignis_b200.device import RAY_DTYPE, Runtime
    t = scene("diamond_scene.json")
    rng = np.random.default_rng(5)
    rays = np.zeros(4096, RAY_DTYPE)
    rays["org"] = rng.uniform(t.bbox_min, t.bbox_max, (4096, 3))
    d = rng.normal(size=(4096, 3))
    rays["dir"] = d * rng.uniform(0.5, 2.0, (4096, 1))   # unnormalised: the device normalises (Device.cpp:617-640)
    rays["tmin"], rays["tmax"] = 1e-3, 1e30
    with plugin.PluginRuntime(t, 4096, 1, 1, tracer=True) as rt:
        got = rt.trace(rays)
    n = rays.copy()
    dd = rays["dir"].astype(np.float32)
    n["dir"] = dd / np.sqrt((dd * dd).sum(1, dtype=np.float32))[:, None]
    with Runtime(t, 4096, 1, spi=1) as rt2:
        ref = rt2.trace(n)
    assert np.isfinite(got).all() and ref.sum() > 0
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)


def _visible_gpus():
    n = C.c_int(0)
    from ignis_b200 import device
    try:
        device.lib().igb200_device_count(C.byref(n))
    except Exception:
        return 0
    return n.value


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spi", [("diamond_scene.json", 200, 120, 2), ("evaluation/multilight-hierarchy.json", 96, 96, 2)])
def test_plugin_device_spreads_the_frame_over_its_gpus(name, w, h, spi, monkeypatch):
    """IGB200_GPUS: one B200Device, several GPUs behind it (tiles + NCCL gather inside the plugin) -- same frame, same counters, AOVs included;
    igtrace's list emitter as well."""
    if _visible_gpus() < 2:
        pytest.skip("needs two GPUs")
    from oracle.oracle import Oracle
    from ignis_b200.device import RAY_DTYPE
    monkeypatch.setenv("IGB200_GPUS", "2")
    t = scene(name)
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(3):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    with plugin.PluginRuntime(t, w, h, spi) as rt:
        assert rt.gpuCount() == 2
        rt.step()
        rt.step()
        first = rt.getFramebufferForHost().copy()      # a gather in the middle of the run must not disturb it
        rt.step()
        got = rt.getFramebufferForHost().copy()
        normals = rt.getFramebufferForHost("Normals").copy()
        stats = rt.getStatistics()
    assert np.isfinite(first).all() and first.sum() > 0
    err = float(np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel()))
    assert err <= 1e-4, err
    assert (stats["CameraRayCount"], stats["ShadowRayCount"], stats["BounceRayCount"]) == tuple(int(x) for x in o.counters)
    monkeypatch.setenv("IGB200_GPUS", "1")
    with plugin.PluginRuntime(t, w, h, spi) as rt1:
        assert rt1.gpuCount() == 1
        rt1.step()
        n1 = rt1.getFramebufferForHost("Normals").copy()
    np.testing.assert_allclose(normals, n1, rtol=1e-5, atol=1e-6)   # written at iteration 0 only
    # igtrace through two GPUs
    monkeypatch.setenv("IGB200_GPUS", "2")
    rng = np.random.default_rng(7)
    rays = np.zeros(4096, RAY_DTYPE)
    rays["org"] = rng.uniform(t.bbox_min, t.bbox_max, (4096, 3))
    rays["dir"] = rng.normal(size=(4096, 3))
    rays["tmin"], rays["tmax"] = 1e-3, 1e30
    with plugin.PluginRuntime(t, 4096, 1, 1, tracer=True) as rt2:
        two = rt2.trace(rays)
    monkeypatch.setenv("IGB200_GPUS", "1")
    with plugin.PluginRuntime(t, 4096, 1, 1, tracer=True) as rt3:
        one = rt3.trace(rays)
    np.testing.assert_allclose(two, one, rtol=1e-4, atol=1e-5)


def test_recogniser_reads_the_other_forms_the_generators_emit(tmp_path):
    """Forms the reference's generators produce that refscript does not (it only has the loader's final values to go by): the roughness +
    anisotropic distribution block of BSDF::setupRoughness (BSDF.cpp:88-96) and a light table that mixes an embedded class with lights
    written into the text (LoaderLight.cpp:199-246)."""
    import re
    t = scene("<procedural>")
    st = refscript.generate(t, cache_dir=str(tmp_path))
    g = plugin.Params(st.global_registry)
    # ---- material 1 is the conductor with roughness 0.2, anisotropic 0.4
    i = next(k for k in range(len(t.materials)) if int(t.materials[k]["distribution"]) == 1 and int(t.materials[k]["map_kind"]) == 0 and int(t.materials[k]["tex"][0]) < 0)
    hit = st.hits[i]
    md = re.search(r"let md_(\w+) = @\|ctx : ShadingContext\| microfacet::make_vndf_ggx_distribution\(ctx\.surf\.face_normal, ctx\.surf\.local, [^;]*\);", hit.script)
    assert md, hit.script
    block = (f"let md_{md.group(1)} = @|ctx : ShadingContext| {{ let (ru, rv) = microfacet::compute_explicit(var_num_x_roughness, var_num_x_anisotropic);"
             "microfacet::make_vndf_ggx_distribution(ctx.surf.face_normal, ctx.surf.local, ru, rv) };")
    script = hit.script.replace(md.group(0), '  let var_num_x_roughness = registry::get_local_parameter_f32("x_roughness", 0);\n'
                                             '  let var_num_x_anisotropic = registry::get_local_parameter_f32("x_anisotropic", 0);\n  ' + block)
    local = refscript.Registry(ints=dict(hit.local.ints), floats=dict(hit.local.floats, x_roughness=0.2, x_anisotropic=0.4), vectors=dict(hit.local.vectors), colors=dict(hit.local.colors))
    m = plugin.CompiledStage(refscript.Stage(hit.function, script, local)).material_tex(g, plugin.TextureTable())
    for f in ("alpha_u", "alpha_v"):
        assert np.float32(m[f]).view(np.uint32) == np.float32(t.materials[i][f]).view(np.uint32), (f, m[f], t.materials[i][f])
    assert int(m["distribution"]) == 1
    # an isotropic literal: `?anisotropic && anisotropic == 0` -> aspect 1 exactly
    m0 = plugin.CompiledStage(refscript.Stage(hit.function, hit.script.replace(md.group(0), "  " + block.replace("var_num_x_roughness", "0.250000").replace("var_num_x_anisotropic", "0.000000")), hit.local)).material_tex(g, plugin.TextureTable())
    assert float(m0["alpha_u"]) == float(m0["alpha_v"]) == 0.25
    # the plain ggx / beckmann distributions are not what this device implements
    with pytest.raises(plugin.DeviceError, match="not supported"):
        plugin.CompiledStage(refscript.Stage(hit.function, hit.script.replace("microfacet::make_vndf_ggx_distribution(ctx.surf.face_normal, ", "microfacet::make_ggx_distribution("), hit.local)).material_tex(g, plugin.TextureTable())
    # ---- a table of 12 embedded point lights followed by one written-out point light
    miss = st.miss
    assert "let finite_lights = e_simplepointlight;" in miss.script
    mixed = miss.script.replace("  let finite_lights = e_simplepointlight;\n",
                                "  let light_99 = make_point_light(12, make_vec3(1.000000, 0.000000, 1.000000), make_color(0.000000, 1.000000, 0.000000, 1));\n"
                                "  let finite_lights = LightTable {\n    count = 13,\n    get   = @|id:i32| {\n    if id < 12 {\n      e_simplepointlight.get(id - 0)\n    }\n"
                                "    else {\n    match(id) {\n      12 => light_99,\n      _ => make_null_light(id)\n    }\n    }\n  }};\n")
    db = plugin.FixTableDB(st.fix_tables)
    inf, fin = plugin.CompiledStage(refscript.Stage(miss.function, mixed, miss.local)).lights(g, db)
    assert len(fin) == 13 and [int(x["type"]) for x in fin] == [1] * 13
    for k in range(12):
        np.testing.assert_array_equal(fin[k]["p"].view(np.uint32), t.finite_lights[k]["p"].view(np.uint32))
    np.testing.assert_array_equal(fin[12]["p"][:6], [1, 0, 1, 0, 1, 0])
    # without the scene database the embedded entries cannot be resolved, and that is an error, not a guess
    with pytest.raises(plugin.DeviceError, match="scene database"):
        plugin.CompiledStage(miss).lights(g)
    # image textures name a file through the resource map: an id that is not in it is reported
    tex_hit = next(h for h in st.hits if "make_checkerboard_texture" in h.script)
    bad = re.sub(r"make_checkerboard_texture\([^;]*\);", 'make_image_texture(make_repeat_border(), make_bilinear_filter(), device.load_image_by_id(0, 4), mat3x3_identity());', tex_hit.script, count=1)
    with pytest.raises(plugin.DeviceError, match="resource map"):
        plugin.CompiledStage(refscript.Stage(tex_hit.function, bad, tex_hit.local)).material_tex(g, plugin.TextureTable())


def _perturb(script: str, rng) -> str:
    """The same program with different trivia: blank lines, comments, extra statements that bind nothing, spaces around punctuation outside
    string literals, and the independent registry look-ups (`let var_* = registry::get_*(...)`) of each run of them in another order."""
    lines = script.split("\n")
    out, run = [], []

    def flush():
        if run:
            rng.shuffle(run)
            out.extend(run)
            run.clear()
    for ln in lines:
        if ln.lstrip().startswith("let var_") and "registry::get_" in ln and ln.rstrip().endswith(";"):
            run.append(ln)
            continue
        flush()
        out.append(ln)
    flush()
    res = []
    for ln in out:
        r = rng.random()
        if r < 0.15:
            res.append('  // a comment; with a semicolon, let x = make_point_light(0, a, b); a brace } and a quote " inside')
        elif r < 0.25:
            res.append("")
        elif r < 0.32 and ln.rstrip().endswith(";") and ln.startswith("  let "):
            res.append("  maybe_unused(settings);")
        new, in_str = "", False
        for ch in ln:
            if ch == '"':
                in_str = not in_str
            if not in_str and ch in ",(" and rng.random() < 0.3:
                new += ch + " " * int(rng.integers(1, 3))
            elif not in_str and ch == ")" and rng.random() < 0.2:
                new += " )"
            elif not in_str and ch == " " and rng.random() < 0.1:
                new += "  "
            else:
                new += ch
        res.append(new + (" " * int(rng.integers(0, 3))))
    return "\n".join(res)


@pytest.mark.parametrize("name", ["diamond_scene.json", "evaluation/multilight-hierarchy.json", "<spot>", "<distant>", "<procedural>", "<bitmaps>", "many_point_lights.json", "evaluation/env4k-conditional.json", "evaluation/sphere-light-pure.json"])
def test_recogniser_does_not_depend_on_trivia(name, tmp_path):
    """the stage text has only ever come from this repository's reconstruction of the generators, so at least the
    recogniser must not depend on how that text is laid out: whitespace, comments, no-op statements and the order of independent bindings."""
    t = scene(name)
    st = refscript.generate(t, cache_dir=str(tmp_path))
    g = plugin.Params(st.global_registry)
    db = plugin.FixTableDB(st.fix_tables)
    rng = np.random.default_rng(11)

    def describe(stages):
        tex = plugin.TextureTable(stages.resource_map)
        hits = [plugin.CompiledStage(s) for s in stages.hits]
        miss = plugin.CompiledStage(stages.miss)
        inf, fin = miss.lights(g, db, tex)
        mats = [h.material_tex(g, tex).tobytes() for h in hits]
        tech = miss.technique(g)
        cam = plugin.CompiledStage(stages.raygen).camera(g)
        return mats, tex.records().tobytes(), tex.aux().tobytes(), inf.tobytes(), fin.tobytes(), tech.tobytes(), miss.selector_data.tobytes(), cam.tobytes()
    ref = describe(st)
    for _ in range(4):
        pert = refscript.StageSet(refscript.Stage(st.raygen.function, _perturb(st.raygen.script, rng), st.raygen.local),
                                  refscript.Stage(st.miss.function, _perturb(st.miss.script, rng), st.miss.local),
                                  [refscript.Stage(h.function, _perturb(h.script, rng), h.local) for h in st.hits], st.global_registry, st.fix_tables, st.resource_map)
        assert describe(pert) == ref


def test_host_png_reader_and_srgb_table():
    """The host layer's image reader against the loader's (ignis_b200/scene.py read_png, itself checked against zlib): byte_color_to_linear for all
    256 values (Image.cpp:40-51) and files that exercise every scanline filter, stored / fixed / dynamic deflate blocks and split IDAT chunks."""
    from ignis_b200.scene import srgb_byte_to_linear_byte
    lut = np.ctypeslib.as_array(plugin.lib().igbh_srgb_lut(), shape=(256,))
    np.testing.assert_array_equal(lut, srgb_byte_to_linear_byte())
    t = scene("<bitmaps>")
    tab = plugin.TextureTable([f[0] for f in t.image_files])
    st = refscript.generate(t)
    g = plugin.Params(st.global_registry)
    for h in st.hits:
        plugin.CompiledStage(h).material_tex(g, tab)
    assert [f for f, _ in tab.images()] == [f for f, _ in t.images]
    # a file that is not a PNG, and one that does not exist: reported
    bad = os.path.join(os.path.dirname(t.image_files[1][0]), "not_a.png")
    open(bad, "wb").write(b"JFIF....")
    for path, msg in ((bad, "not a PNG"), (bad + ".missing", "cannot open")):
        tab2 = plugin.TextureTable([path] * 8)
        with pytest.raises(plugin.DeviceError, match=msg):
            for h in st.hits:
                plugin.CompiledStage(h).material_tex(g, tab2)


def test_host_exr_reader():
    """csrc/host/image_io.cpp load_float_image against OpenEXR's own decoding (tools/make_exr_fixtures.py): every lossless compression, HALF and
    FLOAT samples, partial last chunks; rows come out bottom-up with alpha 1, as the reference's device keeps a float image. In the build container
    every EXR file of the reference tree (55 files from four different writers, the 4k PIZ environment map included) is decoded as well and compared
    with OpenCV's bundled OpenEXR."""
    gold = os.path.join(ROOT, "tests", "golden", "exr")
    exp = np.load(os.path.join(gold, "expected.npz"))
    assert len(exp.files) == 10   # eight EXR files, two Radiance .hdr files (run-length encoded and flat scanlines)
    for name in exp.files:
        got = plugin.load_float_image(os.path.join(gold, name))
        np.testing.assert_array_equal(got[::-1, :, :3], exp[name], err_msg=name)
        assert np.all(got[:, :, 3] == 1)
    with pytest.raises(plugin.DeviceError, match="not an OpenEXR or Radiance"):
        plugin.load_float_image(os.path.join(ROOT, "scenes", "textures", "bumpmap.png"))
    ref_dir = "/root/reference/scenes"
    if os.path.isdir(ref_dir):
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        cv2 = pytest.importorskip("cv2")
        import glob
        files = sorted(glob.glob(os.path.join(ref_dir, "evaluation", "references", "*.exr")))[::6] + [os.path.join(ref_dir, "textures", "environment", "constant.exr")]
        for f in files:
            ref = cv2.imread(f, cv2.IMREAD_UNCHANGED)
            if ref is None:
                continue
            got = plugin.load_float_image(f)
            np.testing.assert_array_equal(got[::-1, :, :3], ref[:, :, 2::-1], err_msg=f)


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spi", [("many_point_lights.json", 192, 108, 1), ("evaluation/env4k-conditional.json", 96, 96, 2), ("evaluation/env4k-none.json", 96, 96, 2)])
def test_render_of_float_image_scenes_through_cpp_plugin(name, w, h, spi):
    """BASELINE config C5 (sky, embedded point lights, bitmap bump map, checkerboard, rough conductor, hierarchy selector) and the environment-map scenes
    through ig_get_interface(): the sky / environment textures travel as OpenEXR files, the cdfs as buffer files, all named through the resource map.
    The descriptors this path hands igb200_set_scene are the loader's bit for bit (test_recognised_descriptors_equal_loader_descriptors, CPU); here the
    frames: plugin == oracle (run on a B200 with the last seconds of round 2's GPU budget: gpurun r6ad, 12 passed)."""
    from oracle.oracle import Oracle
    t = scene(name)
    o = Oracle(t)
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(2):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
    with plugin.PluginRuntime(t, w, h, spi) as rt:
        rt.step()
        rt.step()
        got = rt.getFramebufferForHost().copy()
    err = float(np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel()))
    assert err <= 1e-4, err
