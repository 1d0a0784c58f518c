"""Pins the oracle's round-2 additions: 1-D / 2-D cdfs, GGX / VNDF microfacets, the rough conductor, textures, bump mapping, textured
environment lights and the sky -- against the reference's own known-answer tests where it has them, and against the mathematics
they implement where it has none.

Reference tests restated: src/tests/artic/test_cdf.art (all eight cases), src/tests/artic/test_microfacet.art (ggx iso / aniso:
pdf = D * cos; vndf sample: sample.pdf = pdf(sample)). The reference has no evaluation scene with `conductor`, `bumpmap` or
`checkerboard` (its cycles-* scenes use `principled`), so those are pinned by energy / normalisation / geometry properties.
Reference image used: ref-env4k-4096.exr (textured environment through the 2-D cdf and sampled uniformly). Examined and rejected:
ref-env-4096.exr -- the scene scales a one-pixel map by 100; the reference's make_environment_light_textured applies `scale` to the
emission seen by BSDF-sampled rays but not to light samples (light/env.art:115-126 against :145-152), so its own algorithm does not
reproduce that image (this oracle, which restates the algorithm, gives 1/43 of it; with scale folded out it gives 2.17 x)."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import pytest

from conftest import flat_scene, furnace_scene
from ignis_b200 import scene as S
from ignis_b200.scene import load_scene
from oracle.oracle import Oracle, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fa(*v):
    return (C.c_float * len(v))(*v)


# ------------------------------------------------------------------------------------------------ cdf (src/tests/artic/test_cdf.art)
CDF_DATA = np.array([0.1, 0.2, 0.4, 0.4, 0.8, 1.0], np.float32)   # construct_cdf_1d_test without the leading 0


def cdf1d(u, data=CDF_DATA):
    out = fa(*([0] * 7))
    lib().igo_cdf1d(data.ctypes.data, len(data), u, out)
    return dict(doff=int(out[0]), dpdf=out[1], coff=int(out[2]), pos=out[3], cpdf=out[4], poff=int(out[5]), ppdf=out[6])


def test_cdf_1d_reference_cases():
    r = cdf1d(0.0)      # test_cdf_1d_sample_disc_u_0, ..._cont_u_0
    assert r["doff"] == 0 and r["coff"] == 0 and r["pos"] == 0 and r["cpdf"] == r["ppdf"]
    assert r["dpdf"] == np.float32(0.1)
    r = cdf1d(1.0)      # ..._disc_u_1, ..._cont_u_1
    assert r["doff"] == 5 and r["coff"] == 5 and r["pos"] == 1 and r["cpdf"] == r["ppdf"]
    r = cdf1d(0.79)     # ..._cont_u_079
    assert r["coff"] == 4 and r["cpdf"] == r["ppdf"]
    r = cdf1d(0.8)      # ..._cont_u_08
    assert r["coff"] == 5 and r["cpdf"] == r["ppdf"]
    r = cdf1d(0.57)     # ..._cont_disc
    assert r["doff"] == r["coff"]


def test_cdf_1d_is_a_density():
    # the continuous pdf integrates to one and sampling inverts the cdf
    data = np.cumsum(np.random.default_rng(1).random(37).astype(np.float32) + 0.01, dtype=np.float32)
    data = (data / data[-1]).astype(np.float32)
    data[-1] = 1
    n = len(data)
    pdfs = np.diff(np.concatenate([[0], data])) * n
    assert float(pdfs.mean()) == pytest.approx(1.0, rel=1e-5)
    for u in np.linspace(0.001, 0.999, 97):
        r = cdf1d(float(u), data)
        full = np.concatenate([[0], data])
        assert full[r["coff"]] <= u <= full[r["coff"] + 1] + 1e-7
        assert r["cpdf"] == pytest.approx(pdfs[r["coff"]], rel=1e-5)
        assert r["coff"] / n - 1e-6 <= r["pos"] <= (r["coff"] + 1) / n + 1e-6


def test_cdf_2d_for_image_and_sampling():
    # CDF::computeForImage on a small image: sampling frequencies follow sin-weighted luminance, pdf_continuous agrees with the sample's pdf
    rng = np.random.default_rng(3)
    img = rng.random((6, 9, 3)).astype(np.float32) ** 3
    buf = S.cdf2d_for_image(img, premultiply_sin=True, compensate=False)
    assert buf.shape == (6 + 6 * 9,) and buf[5] == 1 and (np.diff(buf[:6]) >= 0).all()
    us = rng.random((20000, 2)).astype(np.float32)
    hist = np.zeros((6, 9))
    out = fa(0, 0, 0, 0)
    for ux, uy in us:
        lib().igo_cdf2d(buf.ctypes.data, 9, 6, float(ux), float(uy), out)
        assert out[2] == pytest.approx(out[3], rel=1e-5)
        hist[min(int(out[1] * 6), 5), min(int(out[0] * 9), 8)] += 1
    lum = img.mean(axis=2) * np.sin(np.pi * (np.arange(6) + 0.5) / 6)[:, None]
    np.testing.assert_allclose(hist / hist.sum(), lum / lum.sum(), atol=0.006)
    # MIS compensation (CDF.cpp:51-68): a constant image is left alone, else the mean response is subtracted before the cdf is built
    flat = np.full((4, 4, 3), 0.5, np.float32)
    np.testing.assert_array_equal(S.cdf2d_for_image(flat, False, True), S.cdf2d_for_image(flat, False, False))
    comp = S.cdf2d_for_image(img, False, True)
    assert not np.array_equal(comp, S.cdf2d_for_image(img, False, False))


# ------------------------------------------------------------------------------------------------ microfacets (src/tests/artic/test_microfacet.art)
def mf(fn, au, av, w=(0, 0, 1), m=(0, 0, 1), seed=0, counter=1):
    out = fa(0, 0, 0, 0)
    lib().igo_microfacet(fn, au, av, fa(*w), fa(*m), seed, counter, out)
    return tuple(out)


def _setup():   # make_setup, test_microfacet.art:1-6
    wi = np.array([1, 1, 1]) / math.sqrt(3)
    wo = np.array([-1, 1, 1]) / math.sqrt(3)
    h = (wi + wo) / np.linalg.norm(wi + wo)
    return wi, wo, h


@pytest.mark.parametrize("au,av", [(0.1, 0.1), (0.05, 0.45)])
def test_ggx_ndf_closed_form_and_normalisation(au, av):
    wi, wo, h = _setup()
    d = mf(0, au, av, m=h)[0]
    k = (h[0] / au) ** 2 + (h[1] / av) ** 2 + h[2] ** 2
    assert d == pytest.approx(1 / (math.pi * au * av * k * k), rel=1e-5)
    # integral of D(m) cos(theta_m) over the hemisphere is one
    rng = np.random.default_rng(5)
    n = 200000
    u1, u2 = rng.random(n), rng.random(n)
    # importance sample ~ GGX with alpha = max to keep the variance low: plain uniform hemisphere is fine at this n for alpha 0.45 only
    th = np.arctan(max(au, av) * np.sqrt(u1 / (1 - u1)))
    ph = 2 * np.pi * u2
    m = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    a = max(au, av)
    pdf = 1 / (np.pi * a * a * ((m[:, 0] / a) ** 2 + (m[:, 1] / a) ** 2 + m[:, 2] ** 2) ** 2) * m[:, 2]
    vals = np.array([mf(0, au, av, m=mm)[0] * mm[2] for mm in m[:20000]]) / pdf[:20000]
    assert vals.mean() == pytest.approx(1.0, rel=0.03)


def test_vndf_sample_pdf_consistency_and_visible_normals():
    # test_microfacet_vndf_ggx_sample: the pdf returned with a sample is the pdf evaluated for it. (The reference adds the UNSTRETCHED view
    # vector to the spherical-cap sample, core/microfacet.art:386 -- Dupuy & Benyoub add the stretched one -- so with anisotropic roughness a
    # sampled normal can point below the surface; the BSDF flips it towards the viewer, bsdf/conductor.art:108-109. Restated as written.)
    _, wo, _ = _setup()
    for c in range(1, 200):
        nx, ny, nz, pdf = mf(3, 0.05, 0.45, w=wo, seed=42, counter=c)
        n = np.array([nx, ny, nz])
        assert np.linalg.norm(n) == pytest.approx(1, abs=1e-5)
        assert pdf == pytest.approx(mf(2, 0.05, 0.45, w=wo, m=n)[0], rel=1e-6)
    # at normal incidence stretched and unstretched view coincide: every sampled normal is above the surface
    for c in range(1, 200):
        nx, ny, nz, pdf = mf(3, 0.16, 0.16, w=(0, 0, 1), seed=42, counter=c)
        assert nz > 0


def test_smith_g1_limits():
    assert mf(1, 0.3, 0.3, w=(0, 0, 1))[0] == 1.0                 # normal incidence: no masking
    assert mf(1, 0.3, 0.3, w=(1, 0, 0))[0] == 0.0                 # grazing: fully masked (|cos| <= eps)
    g = [mf(1, 0.3, 0.3, w=(math.sin(t), 0, math.cos(t)))[0] for t in np.linspace(0, 1.5, 20)]
    assert all(a >= b for a, b in zip(g, g[1:])) and 0 < g[-1] < 1


def rough_sample(au, av, n, out_dir, seed, counter):
    out = fa(*([0] * 12))
    lib().igo_rough_conductor_sample(au, av, fa(*n), fa(*out_dir), seed, counter, out)
    return list(out)


@pytest.mark.parametrize("alpha", [0.16, 0.4])
def test_rough_conductor_sample_is_consistent_and_loses_little_energy(alpha):
    """bsdf/conductor.art:78-122: the sample's weight is eval / pdf, its pdf is the BSDF's pdf, the direction is the mirror direction
    about a visible microfacet normal; a lossless conductor (eta 0, k 1 -> Fresnel 1) reflects <= 1 and nearly 1 at low roughness."""
    n = np.array([0.0, 0.0, 1.0])
    for theta in (0.2, 0.9, 1.3):
        wo = np.array([math.sin(theta), 0, math.cos(theta)])
        ws, valid = [], 0
        for c in range(1, 1500):
            r = rough_sample(alpha, alpha, n, wo, 7, 2 * c)
            if r[0] == 0:
                ws.append(0.0)
                continue
            valid += 1
            wi = np.array(r[1:4])
            assert np.linalg.norm(wi) == pytest.approx(1, abs=1e-4)     # (the "two-sided" BSDF also returns directions below the surface: |cos| tests)
            assert r[4] == pytest.approx(r[8], rel=1e-4)                                  # sample.pdf == bsdf.pdf(in, out)
            assert r[5] == pytest.approx(r[9] / r[4], rel=1e-4)                            # colour == eval / pdf
            ws.append(r[5])
        albedo = float(np.mean(ws))
        assert valid > 1000
        assert albedo <= 1.0 + 1e-3
        assert albedo >= (0.9 if alpha < 0.2 else 0.8)


def test_rough_conductor_white_furnace_scene():
    """A rough conductor box (roughness 0.16) under a white environment: single-scattering GGX loses only masked light."""
    s = furnace_scene()
    s["bsdfs"] = [{"type": "conductor", "name": "glass", "roughness": 0.16}]
    t = load_scene(s)
    assert int(t.materials[0]["distribution"]) == S.MICROFACET_VNDF_GGX and float(t.materials[0]["alpha_u"]) == np.float32(0.16)
    o = Oracle(t)
    fb = np.zeros((64, 64, 3), np.float32)
    for it in range(8):
        o.render(64, 64, spi=8, iteration=it, fb=fb)
    m = float((fb / 8).mean())
    assert 0.93 < m <= 1.0 + 2e-3
    # a roughness at or below 1e-4 is the delta distribution (core/microfacet.art:297): exactly the mirror furnace
    s["bsdfs"] = [{"type": "conductor", "name": "glass", "roughness": 0.00005}]
    o = Oracle(load_scene(s))
    fb[:] = 0
    o.render(64, 64, spi=4, iteration=0, fb=fb)
    assert float(fb.mean()) == pytest.approx(1.0, abs=1e-5)


# ------------------------------------------------------------------------------------------------ textures
def _tex_scene(textures, reflectance):
    s = flat_scene()
    s["textures"] = textures
    s["bsdfs"][0]["reflectance"] = reflectance
    s["lights"].append({"type": "env", "name": "_e", "radiance": [1, 1, 1]})
    return s


def test_checkerboard_texture_values():
    t = load_scene(_tex_scene([{"type": "checkerboard", "name": "check", "scale_x": 10, "scale_y": 4, "color0": [0.3, 0.2, 0.1], "color1": [1, 0.9, 0.8]}], "check"))
    assert int(t.materials[0]["tex"][0]) == 0 and int(t.textures[0]["type"]) == S.TEX_CHECKERBOARD
    o = Oracle(t)
    rng = np.random.default_rng(0)
    uv = rng.uniform(-2, 3, (4000, 2)).astype(np.float32)
    got = o.eval_texture(0, uv)
    px = np.floor(uv[:, 0].astype(np.float32) * np.float32(10)).astype(np.int64) % 2 == 0
    py = np.floor(uv[:, 1].astype(np.float32) * np.float32(4)).astype(np.int64) % 2 == 0
    want = np.where((px != py)[:, None], np.float32([0.3, 0.2, 0.1]), np.float32([1, 0.9, 0.8]))
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(S.eval_texture_np(t.textures[0], t.images, uv[:, 0], uv[:, 1]), want)   # the loader's own evaluation (baking)


@pytest.mark.parametrize("filt,wrap", [("nearest", "repeat"), ("bilinear", "clamp"), ("bicubic", "mirror"), ("bilinear", "repeat")])
def test_image_texture_filters_and_borders(tmp_path, filt, wrap):
    rng = np.random.default_rng(4)
    img = rng.random((5, 7, 3)).astype(np.float32)
    np.save(tmp_path / "img.npy", img)
    t = load_scene(_tex_scene([{"type": "image", "name": "im", "filename": str(tmp_path / "img.npy"), "filter_type": filt, "wrap_mode": wrap}], "im"))
    fmt, arr = t.images[0]
    assert fmt == S.IMAGE_RGBA32F and arr.shape == (5, 7, 4)
    np.testing.assert_array_equal(arr[::-1, :, :3], img)          # rows bottom-up on the device
    o = Oracle(t)
    # texel centres return the texel; the row order is flipped: v = 0 is the LAST row of the file
    cx, cy = np.meshgrid((np.arange(7) + 0.5) / 7, (np.arange(5) + 0.5) / 5)
    uv = np.stack([cx.ravel(), cy.ravel()], axis=1).astype(np.float32)
    got = o.eval_texture(0, uv).reshape(5, 7, 3)
    if filt != "bicubic":                                         # the cubic B-spline smooths: it does not interpolate
        np.testing.assert_allclose(got, img[::-1], atol=2e-6)
    # the loader's numpy evaluation (used for baking) is the same function
    uv = rng.uniform(-1.5, 2.5, (3000, 2)).astype(np.float32)
    np.testing.assert_allclose(S.eval_texture_np(t.textures[0], t.images, uv[:, 0], uv[:, 1]), o.eval_texture(0, uv), atol=3e-6)
    # smooth filters are continuous across texel borders, all filters respect the wrap mode
    if wrap == "repeat":
        np.testing.assert_allclose(o.eval_texture(0, uv), o.eval_texture(0, uv + np.float32([1, 2])), atol=3e-5)
    if wrap == "clamp":
        far = np.stack([np.full(50, 7.3), np.linspace(0.1, 0.9, 50)], axis=1).astype(np.float32)
        edge = np.stack([np.full(50, 1.0 - 1e-4), np.linspace(0.1, 0.9, 50)], axis=1).astype(np.float32)
        np.testing.assert_allclose(o.eval_texture(0, far), o.eval_texture(0, edge), atol=2e-3)


def test_png_decoder_and_srgb_bytes():
    px = S.read_png(os.path.join(ROOT, "scenes", "textures", "bumpmap.png"))
    assert px.shape == (500, 500, 4) and px.dtype == np.uint8 and int(px[..., :3].min()) == 7 and int(px.max()) == 255
    lut = S.srgb_byte_to_linear_byte()
    assert lut[0] == 0 and lut[255] == 255 and lut[128] == 55 and (np.diff(lut.astype(int)) >= 0).all()   # floor(((128/255 + .055)/1.055)^2.4 * 255) = 55
    fmt, arr = S.load_image_for_device(os.path.join(ROOT, "scenes", "textures", "bumpmap.png"))
    assert fmt == S.IMAGE_RGBA8 and arr.shape == (500, 500, 4)
    np.testing.assert_array_equal(arr[::-1, :, 0], lut[px[:, :, 0]])
    np.testing.assert_array_equal(arr[::-1, :, 3], px[:, :, 3])


def test_textured_diffuse_furnace():
    """A checkerboard-textured plane under a white environment with max_depth 2 returns the texture itself."""
    s = _tex_scene([{"type": "checkerboard", "name": "check", "scale_x": 2, "scale_y": 2, "color0": [0.25, 0.25, 0.25], "color1": [0.75, 0.75, 0.75]}], "check")
    t = load_scene(s)
    o = Oracle(t)
    fb = np.zeros((64, 64, 3), np.float32)
    for it in range(16):
        o.render(64, 64, spi=8, iteration=it, fb=fb)
    img = fb / 16
    vals = img[8:56:16, 8:56:16].mean(axis=2)
    assert set(np.round(vals.ravel() * 4).astype(int)) == {1, 3}
    assert float(img.mean()) == pytest.approx(0.5, abs=0.02)


# ------------------------------------------------------------------------------------------------ bump / normal mapping
def test_normal_set_frame_geometry():
    rng = np.random.default_rng(2)
    for _ in range(200):
        ng = np.array([0, 0, 1.0])
        ns = ng + rng.normal(0, 0.05, 3)
        ns /= np.linalg.norm(ns)
        d = rng.normal(size=3)
        d[2] = -abs(d[2]) - 0.05
        d /= np.linalg.norm(d)
        nn = ng + rng.normal(0, 0.4, 3)
        nn /= np.linalg.norm(nn)
        out = fa(*([0] * 9))
        lib().igo_normal_set_frame(fa(*ng), fa(*ns), fa(*d), fa(*nn), out)
        m = np.array(list(out)).reshape(3, 3).T          # columns
        np.testing.assert_allclose(m.T @ m, np.eye(3), atol=2e-5)               # still an orthonormal frame
        n_new = m[:, 2]
        # ensure_valid_reflection (core/sampling.art:120-165): the mirror direction about the new normal stays above the surface
        r = 2 * (n_new @ -d) * n_new + d
        assert r @ ng >= min(0.9 * (ng @ -d), 0.01) - 2e-4
        if (2 * (nn @ -d) * nn + d) @ ng >= min(0.9 * (ng @ -d), 0.01):
            np.testing.assert_allclose(n_new, nn, atol=2e-5)                      # a valid normal is kept as it is


def test_bumpmap_flat_map_changes_nothing(tmp_path):
    """A constant height map has zero gradient: make_bumpmap must reproduce the un-mapped BSDF (same frame -> same image bit for bit)."""
    np.save(tmp_path / "flat.npy", np.full((4, 4, 3), 0.5, np.float32))
    base = furnace_scene()
    base["bsdfs"] = [{"type": "conductor", "name": "glass", "roughness": 0.2}]
    base["lights"].append({"type": "point", "name": "p", "position": [3, -2, 4], "intensity": [20, 20, 20]})
    a = Oracle(load_scene(base)).render(48, 48, spi=4)
    bumped = json.loads(json.dumps(base))
    bumped["textures"] = [{"type": "image", "name": "flat", "filename": str(tmp_path / "flat.npy"), "filter_type": "bilinear"}]
    bumped["bsdfs"] = [{"type": "conductor", "name": "inner", "roughness": 0.2}, {"type": "bumpmap", "name": "glass", "bsdf": "inner", "map": "flat", "strength": 0.7}]
    t = load_scene(bumped)
    assert int(t.materials[0]["map_kind"]) == S.MAP_BUMP and int(t.materials[0]["bsdf"]) == S.BSDF_CONDUCTOR
    b = Oracle(t).render(48, 48, spi=4)
    assert np.abs(a - b).max() <= 1e-4 * max(1.0, float(a.max()))


# ------------------------------------------------------------------------------------------------ textured environment, sky, C5
def relmse(img, ref):
    mask = ref != 0
    err = np.zeros_like(ref)
    err[mask] = np.square((img[mask] - ref[mask]) / ref[mask])
    err[~mask] = np.square(img[~mask])
    return float(np.average(np.clip(err, 0, np.percentile(err, 99))))


@pytest.mark.parametrize("name", ["env4k-conditional", "env4k-none"])
def test_environment_map_matches_reference_image(name):
    """scenes/evaluation/env4k-*.json against ref-env4k-4096.exr (RunEvaluations.py eps 2e-3 at 1024 spp). The 93 MB map is stored
    box-filtered to 1024 x 512, so the directly visible background differs texel by texel: the RelMSE rule is applied to the lit floor,
    the background is compared through 32 x 32 block means."""
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_images.npz"))["ref-env4k-4096"].astype(np.float32)
    t = load_scene(os.path.join(ROOT, "scenes", "evaluation", name + ".json"))
    assert int(t.infinite_lights[0]["type"]) == (S.LIGHT_ENV_TEXTURED if name.endswith("conditional") else S.LIGHT_ENV_TEX)
    o = Oracle(t)
    fb = np.zeros((256, 256, 3), np.float32)
    spp = 256
    for it in range(spp // 8):
        o.render(256, 256, spi=8, iteration=it, fb=fb)
    img = fb / (spp // 8)
    floor = (slice(210, 244), slice(64, 192))     # inside the trapezoid the floor covers
    # (uniform sampling of an HDR map is far noisier than sampling through its cdf: the bound scales with the variance, not the bias)
    assert relmse(img[floor], ref[floor]) < 2e-3 * (1024 / spp) * (1.5 if name.endswith("conditional") else 4)
    bm = lambda a: a.mean(axis=2).reshape(8, 32, 8, 32).mean(axis=(1, 3))
    np.testing.assert_allclose(bm(img), bm(ref), rtol=0.08)
    assert float(img.mean()) == pytest.approx(float(ref.mean()), rel=0.01)


def test_textured_environment_nee_is_unbiased():
    """Light sampling through the 2-D cdf and pure BSDF sampling estimate the same image: sample_direct, pdf_direct and emission of
    make_environment_light_textured are consistent (direction <-> uv mapping, sin(theta) Jacobian, cdf lookup)."""
    doc = json.load(open(os.path.join(ROOT, "scenes", "evaluation", "env4k-base.json")))
    doc["lights"] = [{"name": "env", "type": "env", "radiance": "env", "cdf": "conditional"}]
    means = {}
    for nee in (True, False):
        doc["technique"]["nee"] = nee
        t = load_scene(doc, base_dir=os.path.join(ROOT, "scenes", "evaluation"))
        o = Oracle(t)
        fb = np.zeros((64, 64, 3), np.float32)
        n = 8 if nee else 48
        for it in range(n):
            o.render(64, 64, spi=16, iteration=it, fb=fb)
        means[nee] = float((fb / n)[53:60, 16:48].mean())
    assert means[True] == pytest.approx(means[False], rel=0.02) and means[True] > 0.1


def test_sky_fixture_and_c5_scene_loads():
    """BASELINE config C5 (scenes/many_point_lights.json): hierarchy selector over ten embedded point lights + the sky model, a
    checkerboard-textured diffuse ground and a bump-mapped rough conductor."""
    t = load_scene(os.path.join(ROOT, "scenes", "many_point_lights.json"))
    assert t.embedded_lights and len(t.finite_lights) == 10 and int(t.technique["light_selector"]) == S.SELECTOR_HIERARCHY
    assert [int(x) for x in t.infinite_lights["type"]] == [S.LIGHT_ENV_TEXTURED]
    fmt, sky = t.images[int(t.textures[int(t.infinite_lights[0]["p"][12:13].view(np.int32)[0])]["image"])]
    assert fmt == S.IMAGE_RGBA32F and sky.shape == (256, 512, 4) and float(sky[..., :3].min()) >= 0
    # default date and place: sun elevation 52.87 degrees, azimuth 323.27 degrees west of south (skysun/SunLocation.h:41): along the row of
    # that elevation the sky is brightest at the sun's azimuth (column x <-> azimuth 360 deg * x / 512 - 45 deg, SkyModel.cpp:35-38)
    row = 255 - int(round((90 - 52.87) / 90 * 256))     # the device holds the rows bottom-up; file row y <-> theta = 90 deg * y / 256
    col = int(np.argmax(sky[row, :, :3].sum(axis=1)))
    assert (col / 512 * 360 - 45) % 360 == pytest.approx(323.27, abs=2.0)
    m = t.materials
    assert int(m[0]["bsdf"]) == S.BSDF_DIFFUSE and int(m[0]["tex"][0]) >= 0
    assert int(m[1]["bsdf"]) == S.BSDF_CONDUCTOR and int(m[1]["map_kind"]) == S.MAP_BUMP and float(m[1]["map_strength"]) == np.float32(0.2)
    assert float(m[1]["alpha_u"]) == np.float32(0.16) and int(m[1]["distribution"]) == S.MICROFACET_VNDF_GGX
    o = Oracle(t)
    img = o.render(96, 96, spi=4)
    assert np.isfinite(img).all() and 0.05 < float(img.mean()) < 1.0
    # reproducible, and NEE through the hierarchy agrees with the uniform selector in the mean
    np.testing.assert_array_equal(img, Oracle(t).render(96, 96, spi=4))


# ------------------------------------------------------------------------------------------------ deterministic accumulation
def test_oracle_deterministic_accumulation():
    """The oracle's per-sample accumulation (counterpart of the device option "deterministic"): independent of the number of threads
    (tiles are independent anyway) and the same image as the reference-order accumulation up to float reordering; at spi 1 the two orders
    coincide in the first iteration (one chain per pixel, added to an empty frame). src/tests/integrator/test_reproducibility.py:5-11."""
    t = load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json"))
    w, h = 96, 54
    def run(det, spi, threads, iters=2):
        o = Oracle(t); o.set_deterministic(det)
        fb = np.zeros((h, w, 3), np.float32)
        for it in range(iters):
            o.render(w, h, spi=spi, iteration=it, fb=fb, threads=threads)
        return fb
    a, b = run(True, 4, 1), run(True, 4, 5)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    plain = run(False, 4, 3)
    assert np.linalg.norm((a - plain).ravel()) / np.linalg.norm(plain.ravel()) < 1e-6
    np.testing.assert_array_equal(run(True, 1, 2, 1).view(np.uint32), run(False, 1, 2, 1).view(np.uint32))   # first iteration: 0 + x is exact


def test_sky_fixture_is_what_the_reference_sources_bake():
    """scenes/textures/sky/*.npz is the output of the reference's OWN sky-model sources (ArHosekSkyModel.cpp, SunLocation.cpp) compiled where they
    lie (oracle/Makefile: _ref/skybake). In the build container, where /root/reference exists, bake C5's sky again and compare bit for bit."""
    import json
    import subprocess
    if not os.path.isdir("/root/reference/src/runtime/skysun"):
        pytest.skip("the reference tree is not here (GPU box): the committed fixture is what travels")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/skybake"], check=True, capture_output=True)
    lj = next(l for l in json.load(open(os.path.join(ROOT, "scenes", "many_point_lights.json")))["lights"] if l.get("type") == "sky")
    g = lj.get("ground", [0.8, 0.8, 0.8])
    g = [g] * 3 if isinstance(g, (int, float)) else g
    args = [os.path.join(ROOT, "oracle", "_ref", "skybake"), *[str(x) for x in g], str(lj.get("turbidity", 3.0))]
    assert not any(k in lj for k in ("direction", "sun_direction", "elevation", "azimuth"))   # C5's sky uses the default date and place
    args += ["time", *[str(lj.get(k, d)) for k, d in (("year", 2020), ("month", 5), ("day", 6), ("hour", 12), ("minute", 0), ("seconds", 0.0),
                                                       ("latitude", 49.235422), ("longitude", -6.9965744), ("timezone", -2))]]
    rgb = np.frombuffer(subprocess.run(args, check=True, capture_output=True).stdout, np.float32).reshape(256, 512, 3)
    fixture = np.load(os.path.join(S.SKY_DIR, f"sky_{S.sky_key(lj)}.npz"))["rgb"]
    np.testing.assert_array_equal(rgb.view(np.uint32), fixture.view(np.uint32))
