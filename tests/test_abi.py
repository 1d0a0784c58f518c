"""The C-ABI library loads and exports every symbol include/igb200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ignis_b200 import device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "igb200.h")).read()
    return sorted(set(re.findall(r"\b(igb200_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    L = device.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/igb200.h but not exported"
    assert sorted(device.SYMBOLS) == syms


def test_version_matches_reference_runtime():
    major, minor = C.c_int(), C.c_int()
    assert device.lib().igb200_version(C.byref(major), C.byref(minor)) == 0
    assert (major.value, minor.value) == (0, 3)   # CMakeLists.txt:25 of the reference (0.3.x), DeviceManager.cpp:180-194


def test_struct_layouts_match_header():
    from ignis_b200 import scene
    assert C.sizeof(device.LookupEntry) == 16 == scene.LOOKUP_DTYPE.itemsize
    assert C.sizeof(device.CameraDesc) == 56 == scene.CAMERA_DTYPE.itemsize
    assert C.sizeof(device.TechniqueDesc) == 20 == scene.TECHNIQUE_DTYPE.itemsize
    assert C.sizeof(device.Settings) == 32
    assert device.RAY_DTYPE.itemsize == 32 and device.HIT_DTYPE.itemsize == 20
    assert scene.LEAF_DTYPE.itemsize == 96 and scene.MATERIAL_DTYPE.itemsize == 128 and scene.LIGHT_DTYPE.itemsize == 128
    assert scene.TEXTURE_DTYPE.itemsize == 96 and C.sizeof(device.ImageDesc) == 24


def test_struct_layouts_against_the_compiled_header(tmp_path):
    """sizeof / offsetof of every descriptor as a C compiler lays out include/igb200.h == the numpy / ctypes mirrors the host side uses."""
    import subprocess
    from ignis_b200 import scene
    fields = {"igb200_material": ["bsdf", "light_id", "p", "tex", "distribution", "alpha_u", "alpha_v", "map_kind", "map_tex", "map_strength", "reserved"],
              "igb200_texture": ["type", "image", "filter", "border_u", "border_v", "reserved", "transform", "p"],
              "igb200_light": ["type", "entity_id", "p"], "igb200_entity_leaf": ["min", "entity_id", "max", "shape_id", "local", "flags", "mat_id", "user1", "user2"],
              "igb200_camera": ["eye", "dir", "up", "fov", "fov_vertical", "aspect", "tmin", "tmax"],
              "igb200_technique": ["max_depth", "min_depth", "clamp", "nee", "light_selector"],
              "igb200_lookup_entry": ["type_id", "flags", "offset"], "igb200_image": ["format", "width", "height", "reserved", "pixels"],
              "igb200_scene_desc": [f for f, _ in device.SceneDesc._fields_], "igb200_settings": [f for f, _ in device.Settings._fields_]}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "igb200.h")}"', 'int main(void) {']
    for st, fs in fields.items():
        src.append(f'printf("{st} %zu\\n", sizeof({st}));')
        src += [f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));' for f in fs]
    src.append('return 0; }')
    (tmp_path / "l.c").write_text("\n".join(src))
    subprocess.run(["gcc", "-o", str(tmp_path / "l"), str(tmp_path / "l.c")], check=True)
    got = dict(line.split() for line in subprocess.run([str(tmp_path / "l")], check=True, capture_output=True, text=True).stdout.splitlines())
    mirrors = {"igb200_material": scene.MATERIAL_DTYPE, "igb200_texture": scene.TEXTURE_DTYPE, "igb200_light": scene.LIGHT_DTYPE, "igb200_entity_leaf": scene.LEAF_DTYPE,
               "igb200_camera": scene.CAMERA_DTYPE, "igb200_technique": scene.TECHNIQUE_DTYPE, "igb200_lookup_entry": scene.LOOKUP_DTYPE}
    for st, dt in mirrors.items():
        assert int(got[st]) == dt.itemsize, st
        for f in fields[st]:
            assert int(got[f"{st}.{f}"]) == dt.fields[f][1], (st, f)
    for st, ct in (("igb200_scene_desc", device.SceneDesc), ("igb200_settings", device.Settings), ("igb200_image", device.ImageDesc)):
        assert int(got[st]) == C.sizeof(ct), st
        for f in fields[st]:
            assert int(got[f"{st}.{f}"]) == getattr(ct, f).offset, (st, f)


def test_no_silent_cpu_fallback():
    """Without a CUDA device the product path must fail loudly instead of computing on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(device.DeviceError):
        device.B200Device(0)


def test_spi_policy():
    # Runtime.cpp:71-79
    assert device.recommend_spi(1000, 1000, gpu=True) == 8
    assert device.recommend_spi(1920, 1080, gpu=True) == 4
    assert device.recommend_spi(1920, 1080, gpu=False) == 1
    assert device.recommend_spi(256, 256, gpu=True) == 64


def test_device_count_without_a_gpu_is_zero_not_an_error():
    """igb200_device_count is how the plugin device sizes IGB200_GPUS = all: no device is an answer (0), not a failure."""
    import torch
    n = C.c_int(-1)
    assert device.lib().igb200_device_count(C.byref(n)) == 0
    if not torch.cuda.is_available():
        assert n.value == 0
    else:
        assert 0 <= n.value <= torch.cuda.device_count()
