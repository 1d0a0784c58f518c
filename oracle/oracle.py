"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE; see oracle/oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
It takes the same scene tables (`ignis_b200.scene.SceneTables`) the device boundary takes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LookupEntry(C.Structure):
    _fields_ = [("type_id", C.c_uint32), ("flags", C.c_uint32), ("offset", C.c_uint64)]


class CameraDesc(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("dir", C.c_float * 3), ("up", C.c_float * 3), ("fov", C.c_float),
                ("fov_vertical", C.c_int32), ("aspect", C.c_float), ("tmin", C.c_float), ("tmax", C.c_float)]


class TechniqueDesc(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("min_depth", C.c_int32), ("clamp", C.c_float), ("nee", C.c_int32), ("light_selector", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("entities", C.c_void_p), ("n_entities", C.c_int32),
                ("shape_lookups", C.c_void_p), ("n_shapes", C.c_int32),
                ("shape_data", C.c_void_p), ("shape_data_bytes", C.c_uint64),
                ("leaves", C.c_void_p), ("n_leaves", C.c_int32),
                ("entity_per_material", C.c_void_p), ("n_materials", C.c_int32),
                ("materials", C.c_void_p),
                ("infinite_lights", C.c_void_p), ("n_infinite", C.c_int32),
                ("finite_lights", C.c_void_p), ("n_finite", C.c_int32),
                ("camera", CameraDesc), ("technique", TechniqueDesc),
                ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("selector_data", C.c_void_p), ("n_selector_data", C.c_int32),
                ("textures", C.c_void_p), ("n_textures", C.c_int32),
                ("images", C.c_void_p), ("n_images", C.c_int32),
                ("aux_data", C.c_void_p), ("n_aux_data", C.c_int32)]


class ImageDesc(C.Structure):
    _fields_ = [("format", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("reserved", C.c_int32), ("pixels", C.c_void_p)]


class Settings(C.Structure):
    _fields_ = [("device", C.c_int32), ("thread_count", C.c_int32), ("spi", C.c_int32), ("frame", C.c_int32),
                ("iter", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("seed", C.c_int32)]


RAY_DTYPE = np.dtype([("org", "<f4", 3), ("dir", "<f4", 3), ("tmin", "<f4"), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("ent_id", "<i4"), ("prim_id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])


def make_scene_desc(tables):
    """Build a SceneDesc pointing into `tables`' arrays. Returns (desc, keepalive)."""
    keep = [np.ascontiguousarray(tables.entities, np.float32), np.ascontiguousarray(tables.shape_lookups),
            np.ascontiguousarray(tables.shape_data), np.ascontiguousarray(tables.leaves),
            np.ascontiguousarray(tables.entity_per_material, np.int32), np.ascontiguousarray(tables.materials),
            np.ascontiguousarray(tables.infinite_lights), np.ascontiguousarray(tables.finite_lights)]
    d = SceneDesc()
    d.entities, d.n_entities = keep[0].ctypes.data, keep[0].shape[0]
    d.shape_lookups, d.n_shapes = keep[1].ctypes.data, keep[1].shape[0]
    d.shape_data, d.shape_data_bytes = keep[2].ctypes.data, keep[2].nbytes
    d.leaves, d.n_leaves = keep[3].ctypes.data, keep[3].shape[0]
    d.entity_per_material, d.n_materials = keep[4].ctypes.data, keep[4].shape[0]
    d.materials = keep[5].ctypes.data
    d.infinite_lights, d.n_infinite = keep[6].ctypes.data, keep[6].shape[0]
    d.finite_lights, d.n_finite = keep[7].ctypes.data, keep[7].shape[0]
    C.memmove(C.byref(d.camera), tables.camera.tobytes(), C.sizeof(CameraDesc))
    C.memmove(C.byref(d.technique), tables.technique.tobytes(), C.sizeof(TechniqueDesc))
    d.bbox_min[:] = [float(x) for x in tables.bbox_min]
    d.bbox_max[:] = [float(x) for x in tables.bbox_max]
    sel = np.ascontiguousarray(getattr(tables, "selector_data", np.zeros(0, np.float32)), np.float32)
    keep.append(sel)
    d.selector_data, d.n_selector_data = (sel.ctypes.data if sel.size else None), int(sel.size)
    tex = np.ascontiguousarray(getattr(tables, "textures", np.zeros(0, np.uint8)))
    aux = np.ascontiguousarray(getattr(tables, "aux_data", np.zeros(0, np.float32)), np.float32)
    imgs = [(int(f), np.ascontiguousarray(a)) for f, a in getattr(tables, "images", [])]
    img_descs = (ImageDesc * max(len(imgs), 1))()
    for i, (f, a) in enumerate(imgs):
        img_descs[i] = ImageDesc(f, a.shape[1], a.shape[0], 0, a.ctypes.data)
    keep += [tex, aux, imgs, img_descs]
    d.textures, d.n_textures = (tex.ctypes.data if tex.size else None), int(tex.shape[0]) if tex.size else 0
    d.images, d.n_images = (C.cast(img_descs, C.c_void_p).value if imgs else None), len(imgs)
    d.aux_data, d.n_aux_data = (aux.ctypes.data if aux.size else None), int(aux.size)
    return d, keep


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.cpp", "detmath.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.igo_create.restype = C.c_void_p
        L.igo_create.argtypes = [C.POINTER(SceneDesc)]
        L.igo_destroy.argtypes = [C.c_void_p]
        L.igo_set_aovs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.igo_set_deterministic.argtypes = [C.c_void_p, C.c_int]
        L.igo_render.argtypes = [C.c_void_p, C.POINTER(Settings), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.igo_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.igo_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        f3 = C.POINTER(C.c_float)
        L.igo_kat_tri.argtypes = [f3, f3, f3, f3, f3, f3, C.c_float, C.c_float, C.c_int, f3]
        L.igo_kat_box.argtypes = [f3, f3, f3, f3, C.c_float, C.c_float, f3]
        L.igo_random_seed.restype = C.c_uint32
        L.igo_random_seed.argtypes = [C.c_int] * 6
        L.igo_tea.restype = C.c_uint32
        L.igo_tea.argtypes = [C.c_uint32, C.c_uint32]
        L.igo_next_f32.restype = C.c_float
        L.igo_next_f32.argtypes = [C.c_uint32, C.c_uint32]
        L.igo_detmath.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.igo_cosine_hemisphere.argtypes = [C.c_float, C.c_float, f3]
        L.igo_dielectric_sample.argtypes = [C.c_float, C.c_float, f3, f3, C.c_int, C.c_uint32, C.c_uint32, f3]
        L.igo_equal_area_sphere.argtypes = [C.c_float, C.c_float, f3]
        L.igo_warp.argtypes = [C.c_int, f3, f3]
        L.igo_hardware_threads.restype = C.c_int
        L.igo_cdf1d.argtypes = [C.c_void_p, C.c_int, C.c_float, f3]
        L.igo_cdf2d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, f3]
        L.igo_microfacet.argtypes = [C.c_int, C.c_float, C.c_float, f3, f3, C.c_uint32, C.c_uint32, f3]
        L.igo_rough_conductor_sample.argtypes = [C.c_float, C.c_float, f3, f3, C.c_uint32, C.c_uint32, f3]
        L.igo_eval_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.igo_normal_set_frame.argtypes = [f3, f3, f3, f3, f3]
        _LIB = L
    return _LIB


class Oracle:
    """CPU restatement of the reference CPU device for one scene."""

    def __init__(self, tables):
        self.tables = tables
        self._desc, self._keep = make_scene_desc(tables)
        self._h = lib().igo_create(C.byref(self._desc))
        self.counters = np.zeros(3, np.uint64)  # camera, shadow, bounce

    def close(self):
        if self._h:
            lib().igo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_aovs(self, normals, albedo):
        """Normals / Albedo AOVs of the reference's infobuffer wrapper: (H, W, 3) float32 arrays accumulated by render() at iteration 0."""
        for a in (normals, albedo):
            assert a is None or (a.dtype == np.float32 and a.flags.c_contiguous)
        self._aovs = (normals, albedo)
        lib().igo_set_aovs(self._h, None if normals is None else normals.ctypes.data, None if albedo is None else albedo.ctypes.data)

    def set_deterministic(self, on: bool):
        """Per-sample accumulation (every sample sums its contributions in path order, samples are added to the pixel in sample order):
        what the device's option "deterministic" computes, bit for bit."""
        lib().igo_set_deterministic(self._h, 1 if on else 0)

    def render(self, width, height, spi=1, iteration=0, seed=0, frame=0, fb=None, threads=0, use_bvh=True,
               rays=None, partition=(0, 1, 32)):
        """One `render()` iteration accumulated into fb (H, W, 3) float32. use_bvh: 0 / False = brute force, 1 / True = median-split BVH2 (the
        parity walks), 2 = SAH BVH4 with 4-primitive leaves, near child first (what the reference's CPU device walks; the timing arm)."""
        if fb is None:
            fb = np.zeros((height, width, 3), np.float32)
        assert fb.dtype == np.float32 and fb.flags.c_contiguous and fb.size == width * height * 3
        st = Settings(0, 0, spi, frame, iteration, width, height, seed)
        if threads <= 0:
            threads = max(1, lib().igo_hardware_threads())
        cnt = (C.c_uint64 * 3)(0, 0, 0)
        rp = None
        if rays is not None:
            rays = np.ascontiguousarray(rays, RAY_DTYPE)
            rp = rays.ctypes.data
        lib().igo_render(self._h, C.byref(st), rp, fb.ctypes.data, threads, int(use_bvh),
                         partition[0], partition[1], partition[2], cnt)
        self.counters += np.asarray(list(cnt), np.uint64)
        return fb

    def eval_texture(self, tex: int, uv) -> np.ndarray:
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        lib().igo_eval_texture(self._h, tex, uv.ctypes.data, uv.shape[0], out.ctypes.data)
        return out

    def trace_closest(self, rays, flags=None, use_bvh=True):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(rays.shape[0], HIT_DTYPE)
        fl = None if flags is None else np.ascontiguousarray(flags, np.uint32)
        lib().igo_trace_closest(self._h, rays.ctypes.data, None if fl is None else fl.ctypes.data, rays.shape[0],
                                out.ctypes.data, int(use_bvh))
        return out

    def trace_any(self, rays, flags=None, use_bvh=True):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(rays.shape[0], np.int32)
        fl = None if flags is None else np.ascontiguousarray(flags, np.uint32)
        lib().igo_trace_any(self._h, rays.ctypes.data, None if fl is None else fl.ctypes.data, rays.shape[0],
                            out.ctypes.data, int(use_bvh))
        return out


def detmath(fn: str, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(a if b is None else b, np.float32)
    out = np.zeros_like(a)
    lib().igo_detmath({"sin": 0, "cos": 1, "acos": 2, "atan2": 3}[fn], a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size)
    return out
