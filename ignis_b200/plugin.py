"""Drives the C++ plugin classes (csrc/host: B200Device : IRenderDevice, B200CompilerDevice : ICompilerDevice, reached through
`ig_get_interface()`) the way the reference's `IG::Runtime` does: SceneDatabase + registries + stage scripts in, framebuffer out.

`PluginRuntime` is the counterpart of `device.Runtime`; the difference is the boundary it crosses. `device.Runtime` hands the
C ABI ready-made descriptors; `PluginRuntime` hands the plugin what the reference hands its devices -- program text
(`refscript.generate`) -- and the plugin's compiler device recognises the descriptors itself.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from . import refscript
from .device import DeviceError, RAY_DTYPE
from .scene import LIGHT_DTYPE, MATERIAL_DTYPE, CAMERA_DTYPE, TECHNIQUE_DTYPE, SceneTables

_LIB = None
SYMBOLS = ["ig_get_interface", "igbh_last_error", "igbh_interface_version", "igbh_db_create", "igbh_db_destroy", "igbh_db_set_fix", "igbh_db_set_dyn",
           "igbh_db_set_bvh", "igbh_db_set_bbox", "igbh_params_create", "igbh_params_destroy", "igbh_params_set_int", "igbh_params_set_float",
           "igbh_params_set_vec3", "igbh_params_set_color", "igbh_compile", "igbh_describe_material", "igbh_describe_lights", "igbh_describe_technique",
           "igbh_describe_camera", "igbh_set_create", "igbh_set_destroy", "igbh_set_raygen", "igbh_set_miss", "igbh_set_add_hit", "igbh_device_create",
           "igbh_device_destroy", "igbh_device_assign", "igbh_assign_release", "igbh_device_render", "igbh_device_resize", "igbh_device_framebuffer",
           "igbh_device_clear", "igbh_device_stats", "igbh_device_gpu_count", "igbh_textures_create", "igbh_textures_destroy", "igbh_textures_count",
           "igbh_textures_get", "igbh_describe_material_tex", "igbh_describe_lights_db", "igbh_textures_set_resources", "igbh_textures_image_count",
           "igbh_textures_image", "igbh_srgb_lut", "igbh_device_assign_res", "igbh_load_float_image", "igbh_textures_aux", "igbh_describe_lights_tex"]


def lib():
    global _LIB
    if _LIB is None:
        try:
            path = _build.build_host()
        except Exception as e:
            path = _build.HOST_OUT
            if not os.path.exists(path):
                raise DeviceError(f"libigb200_host.so is missing and cannot be built: {e}") from e
        L = C.CDLL(path)
        vp = C.c_void_p
        L.igbh_last_error.restype = C.c_char_p
        L.igbh_interface_version.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for f in ("igbh_db_create", "igbh_params_create", "igbh_set_create"):
            getattr(L, f).restype = vp
        L.igbh_db_destroy.argtypes = [vp]
        L.igbh_db_set_fix.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.c_size_t]
        L.igbh_db_set_dyn.argtypes = [vp, C.c_char_p, vp, C.c_size_t, vp, C.c_size_t]
        L.igbh_db_set_bvh.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.igbh_db_set_bbox.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t]
        L.igbh_params_destroy.argtypes = [vp]
        L.igbh_params_set_int.argtypes = [vp, C.c_char_p, C.c_int]
        L.igbh_params_set_float.argtypes = [vp, C.c_char_p, C.c_float]
        L.igbh_params_set_vec3.argtypes = [vp, C.c_char_p, C.POINTER(C.c_float)]
        L.igbh_params_set_color.argtypes = [vp, C.c_char_p, C.POINTER(C.c_float)]
        L.igbh_compile.restype = vp
        L.igbh_compile.argtypes = [C.c_char_p, C.c_char_p]
        L.igbh_describe_material.argtypes = [vp, vp, vp, vp]
        L.igbh_describe_lights.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int), C.c_int]
        L.igbh_describe_technique.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.igbh_describe_camera.argtypes = [vp, vp, vp, vp]
        L.igbh_set_destroy.argtypes = [vp]
        for f in ("igbh_set_raygen", "igbh_set_miss", "igbh_set_add_hit"):
            getattr(L, f).argtypes = [vp, vp, vp]
        L.igbh_device_create.restype = vp
        L.igbh_device_create.argtypes = [C.c_int]
        L.igbh_device_destroy.argtypes = [vp]
        L.igbh_device_assign.restype = vp
        L.igbh_device_assign.argtypes = [vp, vp, vp, C.c_size_t]
        L.igbh_assign_release.argtypes = [vp]
        L.igbh_device_render.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]
        L.igbh_device_resize.argtypes = [vp, C.c_int, C.c_int]
        L.igbh_device_framebuffer.restype = C.POINTER(C.c_float)
        L.igbh_device_framebuffer.argtypes = [vp, C.c_char_p]
        L.igbh_device_clear.argtypes = [vp]
        L.igbh_device_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.igbh_device_gpu_count.argtypes = [vp]
        L.igbh_textures_create.restype = vp
        L.igbh_textures_destroy.argtypes = [vp]
        L.igbh_textures_count.argtypes = [vp]
        L.igbh_textures_get.argtypes = [vp, C.c_int, vp]
        L.igbh_describe_material_tex.argtypes = [vp, vp, vp, vp, vp]
        L.igbh_textures_set_resources.argtypes = [vp, C.POINTER(C.c_char_p), C.c_int]
        L.igbh_textures_image_count.argtypes = [vp]
        L.igbh_textures_image.restype = C.POINTER(C.c_uint8)
        L.igbh_textures_image.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.igbh_srgb_lut.restype = C.POINTER(C.c_uint8)
        L.igbh_textures_aux.restype = C.POINTER(C.c_float)
        L.igbh_textures_aux.argtypes = [vp, C.POINTER(C.c_size_t)]
        L.igbh_describe_lights_tex.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int), C.c_int]
        L.igbh_load_float_image.restype = C.c_long
        L.igbh_load_float_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), vp, C.c_long]
        L.igbh_device_assign_res.restype = vp
        L.igbh_device_assign_res.argtypes = [vp, vp, vp, C.c_size_t, C.POINTER(C.c_char_p), C.c_int]
        L.igbh_describe_lights_db.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int), C.c_int]
        _LIB = L
    return _LIB


def _err() -> str:
    return lib().igbh_last_error().decode()


class Params:
    """IG::ParameterSet on the C++ side."""

    def __init__(self, reg: refscript.Registry | None = None):
        self.h = lib().igbh_params_create()
        if reg is not None:
            for k, v in reg.ints.items():
                lib().igbh_params_set_int(self.h, k.encode(), int(v))
            for k, v in reg.floats.items():
                lib().igbh_params_set_float(self.h, k.encode(), float(v))
            for k, v in reg.vectors.items():
                lib().igbh_params_set_vec3(self.h, k.encode(), (C.c_float * 3)(*v))
            for k, v in reg.colors.items():
                lib().igbh_params_set_color(self.h, k.encode(), (C.c_float * 4)(*v))

    def close(self):
        if self.h:
            lib().igbh_params_destroy(self.h)
            self.h = None


def load_float_image(path: str) -> np.ndarray:
    """An OpenEXR file decoded by the host layer (csrc/host/image_io.cpp): (H, W, 4) float32, rows bottom-up, as the reference's device keeps it."""
    w, h = C.c_int(), C.c_int()
    n = lib().igbh_load_float_image(path.encode(), C.byref(w), C.byref(h), None, 0)
    if n < 0:
        raise DeviceError(_err())
    out = np.zeros(n, np.float32)
    lib().igbh_load_float_image(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, n)
    return out.reshape(h.value, w.value, 4)


class FixTableDB:
    """A SceneDatabase holding only fix-tables (the embedded light tables of a StageSet): enough for CompiledStage.lights."""

    def __init__(self, fix_tables: dict):
        L = lib()
        self.h = L.igbh_db_create()
        for name, tab in fix_tables.items():
            tab = np.ascontiguousarray(tab, np.float32)
            L.igbh_db_set_fix(self.h, name.encode(), tab.ctypes.data, tab.nbytes, tab.shape[0])

    def close(self):
        if self.h:
            lib().igbh_db_destroy(self.h)
            self.h = None


class TextureTable:
    """The texture table the recogniser builds while the hit stages of a scene are resolved in material order."""

    def __init__(self, resource_map=()):
        self.h = lib().igbh_textures_create()
        if resource_map:   # the files behind the resource ids of the stage text (IRenderDevice::SceneSettings::resource_map)
            arr = (C.c_char_p * len(resource_map))(*[p.encode() for p in resource_map])
            lib().igbh_textures_set_resources(self.h, arr, len(resource_map))

    def images(self):
        """The decoded images, as (format, array) pairs shaped like SceneTables.images."""
        out = []
        for i in range(lib().igbh_textures_image_count(self.h)):
            fmt, w, h, n = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
            p = lib().igbh_textures_image(self.h, i, C.byref(fmt), C.byref(w), C.byref(h), C.byref(n))
            a = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
            if n.value == w.value * h.value * 16:   # RGBA32F
                out.append((fmt.value, a.view(np.float32).reshape(h.value, w.value, 4)))
            else:
                out.append((fmt.value, a.reshape(h.value, w.value) if n.value == w.value * h.value else a.reshape(h.value, w.value, 4)))
        return out

    def aux(self) -> np.ndarray:
        """The words of igb200_scene_desc::aux_data collected so far (2-D cdfs of textured environment lights)."""
        n = C.c_size_t()
        p = lib().igbh_textures_aux(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)

    def records(self) -> np.ndarray:
        from .scene import TEXTURE_DTYPE
        n = lib().igbh_textures_count(self.h)
        out = np.zeros(n, TEXTURE_DTYPE)
        for i in range(n):
            lib().igbh_textures_get(self.h, i, out[i:i + 1].ctypes.data)
        return out

    def close(self):
        if self.h:
            lib().igbh_textures_destroy(self.h)
            self.h = None


class CompiledStage:
    """ICompilerDevice::compileAndGet(script, function) + the stage's LocalRegistry (ShaderOutput<void*>)."""

    def __init__(self, stage: refscript.Stage):
        self.handle = lib().igbh_compile(stage.script.encode(), stage.function.encode())
        if not self.handle:
            raise DeviceError(f"compileAndGet({stage.function}) failed: {_err()}")
        self.local = Params(stage.local)

    def material(self, global_params: Params) -> np.ndarray:
        out = np.zeros((), MATERIAL_DTYPE)
        if lib().igbh_describe_material(self.handle, self.local.h, global_params.h, out.ctypes.data):
            raise DeviceError(_err())
        return out

    def material_tex(self, global_params: Params, textures) -> np.ndarray:
        """As material(), with the scene's texture table (igbh_textures_create) that the stage's textures are entered into."""
        out = np.zeros((), MATERIAL_DTYPE)
        if lib().igbh_describe_material_tex(self.handle, self.local.h, global_params.h, textures.h, out.ctypes.data):
            raise DeviceError(_err())
        return out

    def lights(self, global_params: Params, db=None, textures=None):
        """db: an igbh_db handle holding the embedded light fix-tables, when the stage's finite lights come from them; textures: the scene's
        TextureTable, when environment lights are textured (their texture, image and cdf are entered there)."""
        inf, fin = np.zeros(64, LIGHT_DTYPE), np.zeros(64, LIGHT_DTYPE)
        ni, nf = C.c_int(), C.c_int()
        if lib().igbh_describe_lights_tex(self.handle, self.local.h, global_params.h, db.h if db is not None else None, textures.h if textures is not None else None, inf.ctypes.data, C.byref(ni), fin.ctypes.data, C.byref(nf), 64):
            raise DeviceError(_err())
        return inf[:ni.value].copy(), fin[:nf.value].copy()

    def technique(self, global_params: Params) -> np.ndarray:
        out = np.zeros((), TECHNIQUE_DTYPE)
        sel = np.zeros(1 << 16, np.float32)
        n = C.c_int(0)
        if lib().igbh_describe_technique(self.handle, self.local.h, global_params.h, out.ctypes.data, sel.ctypes.data, sel.size, C.byref(n)):
            raise DeviceError(_err())
        self.selector_data = sel[:min(n.value, sel.size)].copy()   # the buffer the cdf / hierarchy light selector reads
        return out

    def camera(self, global_params: Params) -> np.ndarray:
        out = np.zeros((), CAMERA_DTYPE)
        if lib().igbh_describe_camera(self.handle, self.local.h, global_params.h, out.ctypes.data):
            raise DeviceError(_err())
        return out


class PluginRuntime:
    """What IG::Runtime does with a device plugin (Runtime.cpp:81-142,334-446,532-668), through the C++ classes."""

    def __init__(self, tables: SceneTables, width: int, height: int, spi: int, seed: int = 0, cuda_device: int = 0,
                 specialization: str = "default", tracer: bool = False, std_aovs: bool = True):
        L = lib()
        self.tables, self.width, self.height, self.spi, self.seed = tables, int(width), int(height), int(spi), seed
        import tempfile
        self._cache = tempfile.TemporaryDirectory(prefix="igb200_cache_")   # the loader's cache directory (exported selector buffers)
        self.stages = refscript.generate(tables, specialization=specialization, tracer=tracer, std_aovs=std_aovs, cache_dir=self._cache.name)
        # compileShaders (Runtime.cpp:596-668)
        self.global_params = Params(self.stages.global_registry)
        self.raygen = CompiledStage(self.stages.raygen)
        self.miss = CompiledStage(self.stages.miss)
        self.hits = [CompiledStage(s) for s in self.stages.hits]
        self.set = L.igbh_set_create()
        L.igbh_set_raygen(self.set, self.raygen.handle, self.raygen.local.h)
        L.igbh_set_miss(self.set, self.miss.handle, self.miss.local.h)
        for h in self.hits:
            L.igbh_set_add_hit(self.set, h.handle, h.local.h)
        # setupScene (Runtime.cpp:532-541)
        self.db = L.igbh_db_create()
        ent = np.ascontiguousarray(tables.entities, np.float32)
        L.igbh_db_set_fix(self.db, b"entities", ent.ctypes.data, ent.nbytes, ent.shape[0])
        lk = np.ascontiguousarray(tables.shape_lookups)
        data = np.ascontiguousarray(tables.shape_data)
        L.igbh_db_set_dyn(self.db, b"shapes", lk.ctypes.data, lk.shape[0], data.ctypes.data, data.nbytes)
        leaves = np.ascontiguousarray(tables.leaves)
        types = np.array([int(lk[int(l["shape_id"])]["type_id"]) for l in leaves])
        for prov in (0, 1):   # one scene BVH per shape provider (LoaderEntity.cpp:192-201)
            sel = np.ascontiguousarray(leaves[types == prov])
            if len(sel):
                L.igbh_db_set_bvh(self.db, prov, sel.ctypes.data, sel.nbytes)
        L.igbh_db_set_bbox(self.db, (C.c_float * 3)(*[float(x) for x in tables.bbox_min]), (C.c_float * 3)(*[float(x) for x in tables.bbox_max]), len(self.hits))
        for name, tab in self.stages.fix_tables.items():   # LoaderLight::embedLights (LoaderLight.cpp:397-422)
            tab = np.ascontiguousarray(tab, np.float32)
            L.igbh_db_set_fix(self.db, name.encode(), tab.ctypes.data, tab.nbytes, tab.shape[0])
        self.dev = L.igbh_device_create(cuda_device)
        if not self.dev:
            raise DeviceError(f"createRenderDevice failed: {_err()}")
        epm = np.ascontiguousarray(tables.entity_per_material, np.int32)
        res = self.stages.resource_map
        arr = (C.c_char_p * max(len(res), 1))(*[p.encode() for p in res]) if res else None
        self.keep = L.igbh_device_assign_res(self.dev, self.db, epm.ctypes.data, epm.shape[0], arr, len(res))
        L.igbh_device_resize(self.dev, self.width, self.height)
        self.IterationCount = 0

    def step(self):
        if lib().igbh_device_render(self.dev, self.set, self.global_params.h, self.spi, self.width, self.height, self.IterationCount, 0, self.seed, None, 0):
            raise DeviceError(_err())
        self.IterationCount += 1

    def trace(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        flat = np.ascontiguousarray(rays.view(np.float32).reshape(-1, 8))
        if lib().igbh_device_render(self.dev, self.set, self.global_params.h, 1, len(rays), 1, self.IterationCount, 0, self.seed, flat.ctypes.data, len(rays)):
            raise DeviceError(_err())
        self.IterationCount += 1
        p = lib().igbh_device_framebuffer(self.dev, b"")
        return np.ctypeslib.as_array(p, shape=(len(rays), 3)).copy()

    def getFramebufferForHost(self, name: str = "") -> np.ndarray:
        p = lib().igbh_device_framebuffer(self.dev, name.encode())
        if not p:
            raise DeviceError(_err())
        return np.ctypeslib.as_array(p, shape=(self.height, self.width, 3))

    def gpuCount(self) -> int:
        """GPUs behind this one device (environment IGB200_GPUS, b200_device.cpp)."""
        return int(lib().igbh_device_gpu_count(self.dev))

    def getStatistics(self):
        out = (C.c_uint64 * 3)()
        if lib().igbh_device_stats(self.dev, out):
            raise DeviceError("getStatistics failed")
        return {"CameraRayCount": int(out[0]), "ShadowRayCount": int(out[1]), "BounceRayCount": int(out[2])}

    def close(self):
        L = lib()
        if getattr(self, "_cache", None):
            self._cache.cleanup()
            self._cache = None
        if getattr(self, "dev", None):
            L.igbh_device_destroy(self.dev)
            L.igbh_assign_release(self.keep)
            self.dev = None
        if getattr(self, "set", None):
            L.igbh_set_destroy(self.set)
            self.set = None
        if getattr(self, "db", None):
            L.igbh_db_destroy(self.db)
            self.db = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
