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
          "evaluation/two-planes-mirror.json", "evaluation/sun-on-plane.json", "<spot>", "<distant>"]


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
    for i, h in enumerate(hits):
        m = h.material(g)
        assert int(m["bsdf"]) == int(t.materials[i]["bsdf"]) and int(m["light_id"]) == int(t.materials[i]["light_id"])
        ref_p = t.materials[i]["p"].copy()
        if int(m["bsdf"]) == 2 and mode == "disable":
            ref_p[9] = 0   # eta / k come from the registry, Artic's `?eta` is false: the reference itself takes make_pure_conductor_bsdf then
        if exact:
            np.testing.assert_array_equal(m["p"].view(np.uint32), ref_p.view(np.uint32))
        else:
            np.testing.assert_allclose(m["p"], ref_p, atol=1e-6)
    for stage in [plugin.CompiledStage(st.miss)] + hits[:1]:
        inf, fin = stage.lights(g)
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
                                           ("evaluation/multilight-hierarchy.json", 96, 96, 2), ("evaluation/multilight-simple.json", 96, 96, 2)])
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
