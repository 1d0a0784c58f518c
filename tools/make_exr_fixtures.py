"""Writes the OpenEXR (and two Radiance .hdr) fixtures of tests/golden/exr/ with OpenCV's bundled OpenEXR (every lossless compression, HALF and FLOAT samples, sizes that
make partial last chunks) and stores what OpenEXR itself decodes from them next to them (expected.npz). tests/test_plugin_host.py::
test_host_exr_reader checks csrc/host/image_io.cpp against these without needing OpenCV. Run in the build container: python tools/make_exr_fixtures.py"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "exr")
os.makedirs(OUT, exist_ok=True)
rng = np.random.default_rng(5)
expected = {}
for k, (comp, ty, h, w, kind) in enumerate([("PIZ", "HALF", 45, 70, "smooth"), ("PIZ", "FLOAT", 37, 50, "noise"), ("PIZ", "HALF", 33, 64, "noise"), ("ZIP", "HALF", 40, 31, "smooth"),
                                            ("ZIP", "FLOAT", 17, 23, "noise"), ("ZIPS", "FLOAT", 9, 30, "smooth"), ("RLE", "HALF", 12, 40, "flat"), ("NO", "FLOAT", 5, 7, "noise")]):
    if kind == "noise":
        a = (rng.random((h, w, 3)) * 4).astype(np.float32)
    elif kind == "flat":
        a = np.full((h, w, 3), 0.25, np.float32); a[3:6, 5:20] = 2.0
    else:
        a = np.stack([np.outer(np.linspace(0, 3, h), np.linspace(0.5, 2, w)) + 0.1 * c for c in range(3)], axis=2).astype(np.float32)
    name = f"{k}_{comp.lower()}_{ty.lower()}_{h}x{w}.exr"
    cv2.imwrite(os.path.join(OUT, name), a, [cv2.IMWRITE_EXR_COMPRESSION, getattr(cv2, "IMWRITE_EXR_COMPRESSION_" + comp), cv2.IMWRITE_EXR_TYPE, getattr(cv2, "IMWRITE_EXR_TYPE_" + ty)])
    expected[name] = np.ascontiguousarray(cv2.imread(os.path.join(OUT, name), cv2.IMREAD_UNCHANGED)[:, :, ::-1])   # RGB, rows top-down
for k, (h, w) in enumerate([(20, 33), (9, 7)]):   # Radiance RGBE: run-length encoded scanlines (width >= 8) and flat ones
    a = (rng.random((h, w, 3)) * np.array([0.01, 3, 200])).astype(np.float32)
    a[2:5, 3:6] = 0
    name = f"{8 + k}_radiance_{h}x{w}.hdr"
    cv2.imwrite(os.path.join(OUT, name), a[:, :, ::-1])
    expected[name] = np.ascontiguousarray(cv2.imread(os.path.join(OUT, name), cv2.IMREAD_UNCHANGED)[:, :, ::-1])
np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
print(sorted(expected), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
