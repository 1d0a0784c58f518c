"""Host-side mirror of the reference's render-device interface over the C ABI (include/igb200.h).

`B200Device` has the method set of `IG::IRenderDevice` (src/runtime/device/IRenderDevice.h:14-81) with the same
names and argument meaning; `Runtime` mirrors the part of `IG::Runtime` that drives the hot path
(src/runtime/Runtime.cpp:71-79 SPI policy, :334-387 step, :389-446 trace, :794-833 framebuffer scaling) so that
tests read like the reference's integrator tests. There is no CPU fallback: if the CUDA library is missing or no
sm_100 device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .scene import SceneTables, load_scene

_LIB = None


class DeviceError(RuntimeError):
    pass


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

# every symbol include/igb200.h declares (checked by tests/test_abi.py)
SYMBOLS = ["igb200_last_error", "igb200_version", "igb200_device_count", "igb200_create", "igb200_destroy", "igb200_set_scene", "igb200_resize",
           "igb200_set_partition", "igb200_render", "igb200_sync", "igb200_framebuffer", "igb200_framebuffer_device", "igb200_clear",
           "igb200_upload_framebuffer", "igb200_stats", "igb200_reset_stats", "igb200_kernel_times", "igb200_launch_profile", "igb200_turn_log", "igb200_step_stats", "igb200_set_option",
           "igb200_stream", "igb200_trace_closest", "igb200_trace_any", "igb200_bench_trace", "igb200_test_detmath",
           "igb200_comm_unique_id", "igb200_comm_init", "igb200_comm_gather_framebuffer", "igb200_comm_destroy",
           "igb200_frame_stream_begin", "igb200_frame_stream_share", "igb200_frame_stream_next", "igb200_frame_stream_end", "igb200_set_cache_dir", "igb200_scene_build_info", "igb200_test_bvh_build"]


def test_bvh_build(boxes, builder=0, cache_dir=None, device=None):
    """igb200_test_bvh_build: builds + validates a BVH8 over (n, 6) float32 boxes; builder 0 = host (no GPU needed), 1 = GPU (needs a B200Device)."""
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    out = (C.c_int64 * 4)()
    _check(lib().igb200_test_bvh_build(device._h if device is not None else None, boxes.ctypes.data, boxes.shape[0], int(builder),
                                       None if cache_dir is None else str(cache_dir).encode(), out))
    return dict(zip(("nodes", "levels", "leaves", "sah_x1000"), (int(x) for x in out)))


def library_path() -> str:
    return _build.OUT


def lib():
    """Loads libigb200.so (building it with nvcc if the in-tree binary is missing or stale)."""
    global _LIB
    if _LIB is None:
        try:
            path = _build.build()
        except Exception as e:  # no nvcc: use the prebuilt in-tree library as is
            path = _build.OUT
            if not os.path.exists(path):
                raise DeviceError(f"libigb200.so is missing and cannot be built: {e}") from e
        L = C.CDLL(path)
        L.igb200_last_error.restype = C.c_char_p
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.igb200_version.argtypes = [ip, ip]
        L.igb200_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.igb200_destroy.argtypes = [vp]
        L.igb200_set_scene.argtypes = [vp, C.POINTER(SceneDesc)]
        L.igb200_resize.argtypes = [vp, C.c_int, C.c_int]
        L.igb200_set_partition.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.igb200_render.argtypes = [vp, C.POINTER(Settings), vp, C.c_size_t]
        L.igb200_sync.argtypes = [vp]
        L.igb200_framebuffer.argtypes = [vp, C.c_char_p, C.POINTER(C.POINTER(C.c_float))]
        L.igb200_framebuffer_device.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.igb200_clear.argtypes = [vp, C.c_char_p]
        L.igb200_upload_framebuffer.argtypes = [vp, C.c_char_p, vp]
        L.igb200_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.igb200_reset_stats.argtypes = [vp]
        L.igb200_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.igb200_launch_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.igb200_turn_log.argtypes = [vp, vp, vp, vp, C.c_int, ip]
        L.igb200_step_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.igb200_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.igb200_stream.argtypes = [vp, C.POINTER(vp)]
        L.igb200_set_cache_dir.argtypes = [vp, C.c_char_p]
        L.igb200_scene_build_info.argtypes = [vp, C.POINTER(C.c_int64)]
        L.igb200_test_bvh_build.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_char_p, C.POINTER(C.c_int64)]
        L.igb200_trace_closest.argtypes = [vp, vp, vp, C.c_size_t, vp]
        L.igb200_trace_any.argtypes = [vp, vp, C.c_size_t, vp]
        L.igb200_bench_trace.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.igb200_test_detmath.argtypes = [vp, C.c_int, vp, vp, vp, C.c_size_t]
        L.igb200_comm_unique_id.argtypes = [vp]
        L.igb200_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.igb200_comm_gather_framebuffer.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(C.POINTER(C.c_float))]
        L.igb200_comm_destroy.argtypes = [vp]
        L.igb200_frame_stream_begin.argtypes = [vp, C.c_int]
        L.igb200_frame_stream_next.argtypes = [vp, C.c_int, ip, C.POINTER(C.POINTER(C.c_float))]
        L.igb200_frame_stream_end.argtypes = [vp]
        _LIB = L
    return _LIB


def _check(rc: int) -> None:
    if rc != 0:
        raise DeviceError(f"igb200 error {rc}: {lib().igb200_last_error().decode()}")


def make_scene_desc(tables: SceneTables):
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


class B200Device:
    """IRenderDevice over the C ABI. Method names follow src/runtime/device/IRenderDevice.h."""

    def __init__(self, cuda_device: int = 0):
        self._h = C.c_void_p()
        _check(lib().igb200_create(cuda_device, C.byref(self._h)))
        self._w = self._h_ = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().igb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- IRenderDevice
    def assignScene(self, tables: SceneTables):
        desc, keep = make_scene_desc(tables)
        _check(lib().igb200_set_scene(self._h, C.byref(desc)))
        del keep

    def resize(self, width: int, height: int):
        _check(lib().igb200_resize(self._h, width, height))
        self._w, self._h_ = width, height

    def framebufferWidth(self):
        return self._w

    def framebufferHeight(self):
        return self._h_

    def render(self, spi: int, width: int, height: int, iteration: int, frame: int = 0, user_seed: int = 0, rays=None):
        st = Settings(0, 0, spi, frame, iteration, width, height, user_seed)
        if rays is None:
            _check(lib().igb200_render(self._h, C.byref(st), None, 0))
            self._w, self._h_ = width, height
        else:
            rays = np.ascontiguousarray(rays, RAY_DTYPE)
            _check(lib().igb200_render(self._h, C.byref(st), rays.ctypes.data, rays.shape[0]))
            self._w, self._h_ = rays.shape[0], 1

    def sync(self):
        """Finishes every outstanding path of earlier render() calls (deferred tail) and waits for the device."""
        _check(lib().igb200_sync(self._h))

    def getFramebufferForHost(self, name: str = "") -> np.ndarray:
        """Borrowed view (H, W, 3) of the context-owned host buffer; valid until the next resize."""
        p = C.POINTER(C.c_float)()
        _check(lib().igb200_framebuffer(self._h, name.encode(), C.byref(p)))
        return np.ctypeslib.as_array(p, shape=(self._h_, self._w, 3))

    def getFramebufferForDevice(self, name: str = "") -> int:
        p = C.c_void_p()
        _check(lib().igb200_framebuffer_device(self._h, name.encode(), C.byref(p)))
        return int(p.value)

    def clearFramebuffer(self, name: str = ""):
        _check(lib().igb200_clear(self._h, name.encode()))

    def clearAllFramebuffer(self):
        _check(lib().igb200_clear(self._h, None))

    def syncFramebufferHostToDevice(self, host_rgb: np.ndarray, name: str = ""):
        a = np.ascontiguousarray(host_rgb, np.float32)
        assert a.size == self._w * self._h_ * 3
        _check(lib().igb200_upload_framebuffer(self._h, name.encode(), a.ctypes.data))

    def getStatistics(self):
        out = (C.c_uint64 * 5)()
        ms = C.c_double()
        _check(lib().igb200_stats(self._h, out, C.byref(ms)))
        return {"CameraRayCount": int(out[0]), "ShadowRayCount": int(out[1]), "BounceRayCount": int(out[2]),
                "PrimaryRays": int(out[0] + out[2]), "TotalRays": int(out[0] + out[1] + out[2]),
                "Splats": int(out[3]), "KernelLaunches": int(out[4]), "render_ms": ms.value}

    def resetStatistics(self):
        _check(lib().igb200_reset_stats(self._h))

    # -- device-specific
    def setPartition(self, rank: int, world: int, tile: int = 32):
        _check(lib().igb200_set_partition(self._h, rank, world, tile))

    # -- multi-GPU exchange inside the device (include/igb200.h igb200_comm_*)
    @staticmethod
    def commUniqueId() -> bytes:
        """ncclGetUniqueId: made by rank 0, carried to every rank by the caller."""
        buf = (C.c_uint8 * 128)()
        _check(lib().igb200_comm_unique_id(buf))
        return bytes(buf)

    def commInit(self, rank: int, world: int, unique_id: bytes, tile: int = 32):
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().igb200_comm_init(self._h, rank, world, tile, buf))
        self._rank, self._world = rank, world

    def commGatherFramebuffer(self, name: str = "", to_host: bool = True):
        """The path's one exchange: every rank's tiles -> rank 0. Collective. Rank 0 gets (device pointer, host array or None)."""
        dp, hp = C.c_void_p(), C.POINTER(C.c_float)()
        _check(lib().igb200_comm_gather_framebuffer(self._h, name.encode(), C.byref(dp), C.byref(hp) if to_host else None))
        if not dp.value:
            return None, None
        return int(dp.value), (np.ctypeslib.as_array(hp, shape=(self._h_, self._w, 3)) if to_host else None)

    # -- frame streaming (include/igb200.h igb200_frame_stream_*)
    def frameStreamBegin(self, slots: int = 0):
        _check(lib().igb200_frame_stream_begin(self._h, slots))

    def frameStreamShare(self, key: int):
        """Frames of a multi-rank stream in shared host memory (System V key, the same on every rank; before frameStreamBegin)."""
        _check(lib().igb200_frame_stream_share(self._h, int(key)))

    def frameStreamNext(self, wait: int = 0):
        """(iteration, frame) of the next finished iteration -- a borrowed (H, W, 3) view valid until the next call -- or None."""
        it, p = C.c_int(), C.POINTER(C.c_float)()
        rc = lib().igb200_frame_stream_next(self._h, wait, C.byref(it), C.byref(p))
        if rc < 0:
            _check(rc)
        if rc == 0:
            return None
        return it.value, np.ctypeslib.as_array(p, shape=(self._h_, self._w, 3))

    def frameStreamEnd(self):
        _check(lib().igb200_frame_stream_end(self._h))

    def setOption(self, name: str, value: int):
        _check(lib().igb200_set_option(self._h, name.encode(), int(value)))

    def setCacheDir(self, path):
        """Directory of the on-disk BVH cache (None: no cache); the reference: LoaderContext.CacheManager, TriMeshProvider.cpp:326-351."""
        _check(lib().igb200_set_cache_dir(self._h, None if path is None else str(path).encode()))

    def sceneBuildInfo(self):
        """What the last assignScene did with the shapes' BVHs."""
        out = (C.c_int64 * 6)()
        _check(lib().igb200_scene_build_info(self._h, out))
        return dict(zip(("host_built", "gpu_built", "cache_loaded", "cache_stored", "build_us", "nodes"), (int(x) for x in out)))

    def turnLog(self):
        """Per loop turn of the last render(): (rays traced, trace-phase ns, shade-phase ns)."""
        a, b, c = (np.zeros(128, np.uint32) for _ in range(3))
        n = C.c_int()
        _check(lib().igb200_turn_log(self._h, a.ctypes.data, b.ctypes.data, c.ctypes.data, 128, C.byref(n)))
        return a[:n.value].copy(), b[:n.value].copy(), c[:n.value].copy()

    def stepStats(self):
        out = (C.c_uint64 * 16)()
        _check(lib().igb200_step_stats(self._h, out))
        names = ("node_visits", "leaf_visits", "entity_visits", "max_visits_per_ray", "rays")
        return {"turns<16": {n: int(out[i]) for i, n in enumerate(names)}, "turns>=16": {n: int(out[8 + i]) for i, n in enumerate(names)}}

    def stream(self) -> int:
        p = C.c_void_p()
        _check(lib().igb200_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def kernelTimes(self):
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        _check(lib().igb200_kernel_times(self._h, ms, n))
        return {"trace": {"ms": ms[1], "launches": int(n[1])}, "shade_generate": {"ms": ms[2], "launches": int(n[2])}}

    def launchProfile(self):
        """Per-kernel CUDA-event times (needs setOption("profile_kernels", 1)) and the work of the k_turn_trace launches."""
        ms, n, w = (C.c_double * 4)(), (C.c_uint64 * 4)(), (C.c_uint64 * 3)()
        _check(lib().igb200_launch_profile(self._h, ms, n, w))
        names = ("k_wavefront", "k_turn_trace", "k_turn_shade", "k_turn_end")
        return {"kernels": {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(names)},
                "k_turn_trace_work": {"primary": int(w[0]), "shadow": int(w[1]), "splats": int(w[2])}}

    def traceClosest(self, rays, flags=None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(rays.shape[0], HIT_DTYPE)
        fl = None if flags is None else np.ascontiguousarray(flags, np.uint32)
        _check(lib().igb200_trace_closest(self._h, rays.ctypes.data, None if fl is None else fl.ctypes.data, rays.shape[0], out.ctypes.data))
        return out

    def traceAny(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(rays.shape[0], np.int32)
        _check(lib().igb200_trace_any(self._h, rays.ctypes.data, rays.shape[0], out.ctypes.data))
        return out

    def benchTrace(self, rays, any_hit=False, repeat=10) -> float:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        ms = C.c_double()
        _check(lib().igb200_bench_trace(self._h, rays.ctypes.data, rays.shape[0], 1 if any_hit else 0, repeat, C.byref(ms)))
        return ms.value

    def testDetmath(self, fn: str, a, b=None) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(a if b is None else b, np.float32)
        out = np.zeros_like(a)
        _check(lib().igb200_test_detmath(self._h, {"sin": 0, "cos": 1, "acos": 2, "atan2": 3}[fn], a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size))
        return out


def recommend_spi(width: int, height: int, gpu: bool = True) -> int:
    """Runtime.cpp:71-79: the "best case" was measured with a 1000 x 1000 film."""
    spi_f = 8 if gpu else 2
    spi = int(min(64, max(1.0, np.ceil(spi_f / ((width / 1000.0) * (height / 1000.0))))))
    return spi


class Runtime:
    """The slice of IG::Runtime that drives the device: load, step, trace, framebuffer (sum / iterations)."""

    def __init__(self, scene, width=None, height=None, spi=0, seed=0, cuda_device=0, max_depth=None, device=None):
        self.tables = scene if isinstance(scene, SceneTables) else load_scene(scene, width, height, max_depth)
        self.width = int(width if width is not None else self.tables.film_size[0])
        self.height = int(height if height is not None else self.tables.film_size[1])
        self.spi = int(spi) if spi and spi > 0 else recommend_spi(self.width, self.height, True)
        self.seed = seed
        self.device = device or B200Device(cuda_device)
        self.device.assignScene(self.tables)
        self.device.resize(self.width, self.height)
        self.IterationCount = 0
        self.SampleCount = 0
        self.FrameCount = 0

    def close(self):
        self.device.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self):
        # Runtime::stepVariant, Runtime.cpp:366-387
        self.device.render(self.spi, self.width, self.height, self.IterationCount, self.FrameCount, self.seed)
        self.IterationCount += 1
        self.SampleCount += self.spi

    def trace(self, rays) -> np.ndarray:
        # Runtime::trace, Runtime.cpp:389-446: spi 1, film = rays x 1, returns radiance per ray
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        self.device.clearAllFramebuffer() if self.device.framebufferWidth() else None
        self.device.render(1, rays.shape[0], 1, self.IterationCount, self.FrameCount, self.seed, rays=rays)
        self.IterationCount += 1
        return self.device.getFramebufferForHost().reshape(-1, 3).copy()

    def reset(self):
        self.device.clearAllFramebuffer()
        self.IterationCount = 0
        self.SampleCount = 0

    def getFramebufferForHost(self) -> np.ndarray:
        return self.device.getFramebufferForHost()

    def image(self) -> np.ndarray:
        """Framebuffer scaled by 1 / iterations, as Runtime::saveFramebuffer does (Runtime.cpp:808)."""
        return self.device.getFramebufferForHost() / max(1, self.IterationCount)
