"""Host side of the non-uniform light selectors: the buffers `light_cdf.bin` / `light_hierarchy.bin` the loader
exports (src/runtime/CDF.cpp:14-44, src/runtime/light/LightHierarchy.cpp:47-129, src/runtime/container/PointBvh.inl) as restated in
ignis_b200/scene.py. The estimators that read them are pinned by the reference's converged image of the multilight scene
(tests/test_oracle_images.py: multilight-simple, multilight-hierarchy) and, on the GPU, against the oracle."""
import os

import numpy as np
import pytest

from ignis_b200.scene import SELECTOR_CDF, SELECTOR_HIERARCHY, SELECTOR_UNIFORM, light_cdf, light_hierarchy, load_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cdf_is_the_normalised_running_sum_without_the_leading_zero():
    cdf = light_cdf([1.0, 3.0, 4.0])
    np.testing.assert_allclose(cdf, [0.125, 0.5, 1.0])
    assert cdf.dtype == np.float32 and cdf[-1] == 1.0
    # all-dark lights: uniform (CDF.cpp:31-36)
    np.testing.assert_allclose(light_cdf([0.0, 0.0, 0.0, 0.0]), [0.25, 0.5, 0.75, 1.0])


def _walk(data, n, light):
    """Follows the light's code from the root (light_hierarchy.art:78-95); returns the leaf's light id and the depth."""
    codes = data[:(n + 3) // 4 * 4].view(np.uint32)
    nodes = data[(n + 3) // 4 * 4:].reshape(-1, 8)
    code, k, depth = int(codes[light]), 0, 0
    while nodes[k, 7:8].view(np.int32)[0] < 0:
        left = -int(nodes[k, 7:8].view(np.int32)[0]) - 1
        k = left if (code & 1) == 0 else left + 1
        code >>= 1
        depth += 1
    return int(nodes[k, 7:8].view(np.int32)[0]), depth


@pytest.mark.parametrize("n", [2, 3, 7, 33])
def test_hierarchy_codes_lead_to_their_lights_and_flux_adds_up(n):
    rng = np.random.default_rng(n)
    lights = []
    for i in range(n):
        pos = rng.uniform(-5, 5, 3).astype(np.float32)
        d = rng.normal(size=3)
        lights.append((pos, (d / np.linalg.norm(d)).astype(np.float32) if i % 3 == 0 else None, float(rng.uniform(0.1, 10))))
    data = light_hierarchy(lights)
    nodes = data[(n + 3) // 4 * 4:].reshape(-1, 8)
    assert len(nodes) == 2 * n - 1                                       # a full binary tree over n leaves
    for i in range(n):
        leaf, depth = _walk(data, n, i)
        assert leaf == i and depth <= 32                                 # the codes hold 32 turns (LightHierarchy.cpp:120)
    # a leaf stores the light (flux negative when it has no direction), an inner node the sum of its children's |flux|
    total = sum(f for _, _, f in lights)
    assert abs(nodes[0, 3]) == pytest.approx(total, rel=1e-5)
    for k in range(len(nodes)):
        idx = int(nodes[k, 7:8].view(np.int32)[0])
        if idx >= 0:
            pos, d, f = lights[idx]
            np.testing.assert_array_equal(nodes[k, 0:3], pos)
            assert nodes[k, 3] == pytest.approx(f if d is not None else -f)
        else:
            left = -idx - 1
            assert left > k and left + 1 < len(nodes)                    # children follow their parent: the walk terminates
            assert abs(nodes[k, 3]) == pytest.approx(abs(nodes[left, 3]) + abs(nodes[left + 1, 3]), rel=1e-5)


def test_loader_picks_the_selector_the_reference_would(monkeypatch):
    ev = os.path.join(ROOT, "scenes", "evaluation")
    t = load_scene(os.path.join(ev, "multilight-simple.json"))
    assert int(t.technique["light_selector"]) == SELECTOR_CDF and len(t.selector_data) == len(t.finite_lights) == 3
    # flux (Light::computeFlux): plane light radiance 0.5 * area 4 * pi; point lights mean(intensity) * 4 pi
    np.testing.assert_allclose(np.diff(np.concatenate([[0], t.selector_data])) * (2 * np.pi + 2 * 1.3 / 3 * 4 * np.pi), [2 * np.pi, 1.3 / 3 * 4 * np.pi, 1.3 / 3 * 4 * np.pi], rtol=1e-5)
    t = load_scene(os.path.join(ev, "multilight-hierarchy.json"))
    assert int(t.technique["light_selector"]) == SELECTOR_HIERARCHY and len(t.selector_data) == 4 + 8 * 5
    # one light or none: the generator emits the uniform selector whatever the scene asks for (LoaderLight.cpp:427-429)
    import json
    doc = json.load(open(os.path.join(ev, "point.json")))
    base = json.load(open(os.path.join(ev, doc["externals"][0]["filename"]))) if "externals" in doc else {}
    base.update({k: v for k, v in doc.items() if k != "externals"})
    base.setdefault("technique", {"type": "path"})["light_selector"] = "hierarchy"
    monkeypatch.chdir(ev)   # the scene's mesh paths are relative
    assert int(load_scene(base).technique["light_selector"]) == SELECTOR_UNIFORM


def _many_lights_scene(selector, n=24, seed=5):
    from conftest import flat_scene
    rng = np.random.default_rng(seed)
    s = flat_scene()
    s["technique"]["light_selector"] = selector
    s["film"]["size"] = [64, 64]
    for i in range(n):
        pos = [float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-2.5, -0.3))]
        if i % 4 == 0:
            s["lights"].append({"type": "spot", "name": f"s{i}", "cutoff": 50, "falloff": 35, "position": pos, "direction": [0.1, -0.1, 1], "intensity": [float(rng.uniform(0.2, 3))] * 3})
        else:
            s["lights"].append({"type": "point", "name": f"p{i}", "position": pos, "intensity": [float(x) for x in rng.uniform(0.05, 2.0, 3)]})
    s["lights"].append({"type": "env", "name": "e", "radiance": [0.05, 0.05, 0.05]})
    return s


def test_selectors_are_unbiased_on_many_lights():
    """24 point / spot lights of very different strength plus an environment: the three selectors are different estimators of the same
    integral, so their converged images agree; the hierarchy (importance by flux / distance^2) has the lowest variance."""
    from oracle.oracle import Oracle
    imgs = {}
    for sel in ("uniform", "simple", "hierarchy"):
        t = load_scene(_many_lights_scene(sel))
        assert int(t.technique["light_selector"]) == {"uniform": SELECTOR_UNIFORM, "simple": SELECTOR_CDF, "hierarchy": SELECTOR_HIERARCHY}[sel]
        o = Oracle(t)
        fb = np.zeros((64, 64, 3), np.float32)
        n = 24
        for it in range(n):
            o.render(64, 64, spi=16, iteration=it, fb=fb)
        imgs[sel] = fb / n
    ref = imgs["uniform"]
    for sel in ("simple", "hierarchy"):
        assert imgs[sel].mean() == pytest.approx(ref.mean(), rel=0.01), sel
        # 8x8 block averages: the agreement is local, not just global
        a = imgs[sel].reshape(8, 8, 8, 8, 3).mean(axis=(1, 3)); b = ref.reshape(8, 8, 8, 8, 3).mean(axis=(1, 3))
        assert np.abs(a - b).max() / b.mean() < 0.10, sel   # 384 spp: block noise of the uniform selector alone is ~5 %
