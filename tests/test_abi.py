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
    assert scene.LEAF_DTYPE.itemsize == 96 and scene.MATERIAL_DTYPE.itemsize == 64 and scene.LIGHT_DTYPE.itemsize == 128


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
