"""Bakes the image of a `sky` light (src/runtime/light/SkyLight.cpp:28-47) into scenes/textures/sky/sky_<key>.npz.

Run in the build container, where /root/reference is mounted: compiles the reference's own sky-model sources where they lie
(oracle/Makefile target _ref/skybake) and runs them with the light's parameters. The outputs are committed: the GPU box has no
/root/reference, and ignis_b200/scene.py only ever reads the fixture. Usage: tools/make_sky.py scene.json [...]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignis_b200.scene import SKY_DIR, sky_key  # noqa: E402


def bake(lj: dict) -> str:
    exe = os.path.join(ROOT, "oracle", "_ref", "skybake")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/skybake"], check=True, capture_output=True)
    g = lj.get("ground", [0.8, 0.8, 0.8])
    g = [g] * 3 if isinstance(g, (int, float)) else g
    args = [exe, *[str(x) for x in g], str(lj.get("turbidity", 3.0))]
    if "direction" in lj or "sun_direction" in lj:
        args += ["dir", *[str(x) for x in lj.get("direction", lj.get("sun_direction"))]]
    elif "elevation" in lj or "azimuth" in lj:
        args += ["ea", str(lj.get("elevation", 0)), str(lj.get("azimuth", 0))]
    else:   # SunLocation.h defaults: 2020-05-06 12:00, Saarbruecken
        args += ["time", *[str(lj.get(k, d)) for k, d in (("year", 2020), ("month", 5), ("day", 6), ("hour", 12), ("minute", 0), ("seconds", 0.0),
                                                           ("latitude", 49.235422), ("longitude", -6.9965744), ("timezone", -2))]]
    r = subprocess.run(args, check=True, capture_output=True)
    rgb = np.frombuffer(r.stdout, np.float32).reshape(256, 512, 3)
    os.makedirs(SKY_DIR, exist_ok=True)
    path = os.path.join(SKY_DIR, f"sky_{sky_key(lj)}.npz")
    np.savez_compressed(path, rgb=rgb)
    print(path, r.stderr.decode().strip(), "mean", float(rgb.mean()), "max", float(rgb.max()))
    return path


if __name__ == "__main__":
    for scene in sys.argv[1:]:
        for lj in json.load(open(scene)).get("lights", []):
            if lj.get("type") == "sky":
                bake(lj)
