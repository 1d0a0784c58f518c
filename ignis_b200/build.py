"""Builds ignis_b200/libigb200.so (C ABI + sm_100a kernels) in-tree with nvcc.

-fmad=false: the only fused multiply-adds in device code are explicit (see csrc/device_math.cuh).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libigb200.so")
SOURCES = ["api.cu", "bvh_build.cu"]
DEPS = ["api.cu", "bvh_build.cu", "bvh_build.h", "wavefront.cuh", "traverse.cuh", "shade.cuh", "material.cuh", "types.cuh", "device_math.cuh", "bvh8.h", "../../include/igb200.h", "../build.py"]


def nvcc_path() -> str:
    for p in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    raise RuntimeError("nvcc not found")


def flags(extra=()):
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
            "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-O2", *extra]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(SRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    global OUT
    if os.environ.get("IGB200_EXPERIMENT"):   # kernel experiments: -DIGB_EXP_<name>, separate binary (tools/tune.py)
        OUT = os.path.join(HERE, "libigb200_exp_%s.so" % os.environ["IGB200_EXPERIMENT"])
        force = force or not os.path.exists(OUT)
    if os.environ.get("IGB200_STEP_STATS"):   # diagnostics build: counts BVH visits per ray (slow), separate binary
        OUT = os.path.join(HERE, "libigb200_stats.so")
        force = force or not os.path.exists(OUT)
    if not force and not needs_build():
        return OUT
    extra = (["-Xptxas", "-v"] if verbose else []) + (["-DIGB_STEP_STATS"] if os.environ.get("IGB200_STEP_STATS") else []) + \
            (["-DIGB_EXP_" + x for x in os.environ["IGB200_EXPERIMENT"].split("_")] if os.environ.get("IGB200_EXPERIMENT") else [])
    cmd = [nvcc_path(), *flags(extra), "--threads", "2", "-shared", "-o", OUT,
           *[os.path.join(SRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return OUT


HOST_OUT = os.path.join(HERE, "libigb200_host.so")
HOST_SOURCES = ["host/script_recognizer.cpp", "host/image_io.cpp", "host/b200_device.cpp", "host/host_capi.cpp"]
HOST_DEPS = HOST_SOURCES + ["host/script_recognizer.h", "host/b200_device.h", "host/ig_mirror.h", "../../include/igb200.h"]


def build_host(force: bool = False) -> str:
    """Builds ignis_b200/libigb200_host.so: the C++ plugin classes (IRenderDevice / ICompilerDevice / ig_get_interface)
    over the C ABI, plus the C shim the tests drive them through. Host-only C++17, links libigb200.so."""
    build()
    if not force and os.path.exists(HOST_OUT):
        t = os.path.getmtime(HOST_OUT)
        if not any(os.path.getmtime(os.path.join(SRC, d)) > t for d in HOST_DEPS):
            return HOST_OUT
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", HOST_OUT,
           *[os.path.join(SRC, s) for s in HOST_SOURCES], "-L" + HERE, "-l:libigb200.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return HOST_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
