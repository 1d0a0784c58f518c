"""Scene ingest: Ignis scene JSON -> the binary tables and descriptors the device boundary takes.

This is the caller side of the hot path (what the reference's Loader hands to
`IRenderDevice::assignScene`), restated as data producers:

* `entities` fix-table, 36 f32 per entity           (src/runtime/loader/LoaderEntity.cpp:150-162)
* `shapes` dyn-table: 16-byte lookups + data blob    (src/runtime/shape/TriMeshProvider.cpp:575-596,
                                                      src/runtime/shape/SphereProvider.cpp:42-47,
                                                      src/runtime/table/DynTable.h:6-36)
* scene-BVH leaves `EntityLeaf1`, 96 B per entity    (src/artic/traversal/bvh.art:52-61,
                                                      src/runtime/bvh/SceneBVHAdapter.h:68-100)
* `entity_per_material`                              (src/runtime/loader/LoaderEntity.cpp:42-97)
* material / light / camera / technique descriptors = the arguments of the constructor calls the
  reference's generators emit as shader text (DiffuseBSDF.cpp:13-27, DielectricBSDF.cpp:13-41,
  AreaLight.cpp:115-220, PointLight.cpp:44-62, EnvironmentLight.cpp:44-110,
  PerspectiveCamera.cpp:26-67, PathTechnique.cpp:35-79).

Entity and light order: the reference iterates `std::unordered_map`s (Scene.h:59), i.e. an
implementation-defined order; this loader uses file order.
"""
from __future__ import annotations

import json
import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

from . import meshes
from .meshes import TriMesh

F = np.float32
PI = float(F(3.14159265358979323846))
DEG2RAD = float(F(PI) / F(180.0))

SHAPE_TRIMESH = 0
SHAPE_SPHERE = 1

BSDF_DIFFUSE = 0      # make_lambertian_bsdf (bsdf/diffuse.art:2-12)
BSDF_DIELECTRIC = 1   # make_pure_dielectric_bsdf (bsdf/dielectric.art:15-37)
BSDF_CONDUCTOR = 2    # make_mirror_bsdf / make_pure_conductor_bsdf (bsdf/conductor.art:2-27,131-141): smooth conductors only

LIGHT_ENV_CONST = 0   # make_environment_light -> ..._function_spherical (light/env.art:75-100,161-164)
LIGHT_POINT = 1       # make_point_light (light/point.art:1-18)
LIGHT_PLANE_AREA = 2  # make_area_light + make_plane_area_emitter (light/area.art:10-43,124-258)
LIGHT_SHAPE_AREA = 3  # make_area_light + make_shape_area_emitter (light/area.art:62-107)
LIGHT_SPHERE_AREA = 4  # make_area_light + make_sphere_area_emitter (light/area.art:260-316)
LIGHT_SPOT = 5         # make_spot_light (light/spot.art:8-44)
LIGHT_SUN = 6          # make_sun_light (light/sun.art:10-48): infinite, p = direction towards the sun, cos(angle / 2), radiance
LIGHT_DIRECTIONAL = 7  # make_directional_light (light/directional.art): infinite delta light, p = direction of travel, irradiance
LIGHT_ENV_TEXTURED = 8  # make_environment_light_textured (light/env.art:112-160): env map / sky with a 2-D cdf
LIGHT_ENV_TEX = 9       # make_environment_light over a texture (light/env.art:161-167), sampled uniformly

MICROFACET_DELTA, MICROFACET_VNDF_GGX = 0, 1      # BSDF::setupRoughness (bsdf/BSDF.cpp:53-98), core/microfacet.art:403-425
MAP_NONE, MAP_BUMP, MAP_NORMAL = 0, 1, 2           # bsdf/map.art:56-68
TEX_CHECKERBOARD, TEX_IMAGE = 0, 1                 # texture/checkerboard.art, texture/image.art
FILTERS = {"nearest": 0, "bilinear": 1}            # anything else is the bicubic filter (ImagePattern.cpp:26-30)
FILTER_BICUBIC = 2
BORDERS = {"clamp": 1, "mirror": 2}                # anything else repeats (ImagePattern.cpp:32-39)
IMAGE_RGBA8, IMAGE_MONO8, IMAGE_RGBA32F = 0, 1, 2

LOOKUP_DTYPE = np.dtype([("type_id", "<u4"), ("flags", "<u4"), ("offset", "<u8")])
LEAF_DTYPE = np.dtype([("min", "<f4", 3), ("entity_id", "<i4"), ("max", "<f4", 3), ("shape_id", "<i4"),
                       ("local", "<f4", 12), ("flags", "<u4"), ("mat_id", "<i4"), ("user1", "<i4"), ("user2", "<i4")])
MATERIAL_DTYPE = np.dtype([("bsdf", "<i4"), ("light_id", "<i4"), ("p", "<f4", 14), ("tex", "<i4", 2), ("distribution", "<i4"),
                           ("alpha_u", "<f4"), ("alpha_v", "<f4"), ("map_kind", "<i4"), ("map_tex", "<i4"), ("map_strength", "<f4"), ("reserved", "<i4", 8)])
TEXTURE_DTYPE = np.dtype([("type", "<i4"), ("image", "<i4"), ("filter", "<i4"), ("border_u", "<i4"), ("border_v", "<i4"), ("reserved", "<i4", 3),
                          ("transform", "<f4", 6), ("p", "<f4", 10)])
LIGHT_DTYPE = np.dtype([("type", "<i4"), ("entity_id", "<i4"), ("p", "<f4", 30)])
CAMERA_DTYPE = np.dtype([("eye", "<f4", 3), ("dir", "<f4", 3), ("up", "<f4", 3), ("fov", "<f4"),
                         ("fov_vertical", "<i4"), ("aspect", "<f4"), ("tmin", "<f4"), ("tmax", "<f4")])
TECHNIQUE_DTYPE = np.dtype([("max_depth", "<i4"), ("min_depth", "<i4"), ("clamp", "<f4"), ("nee", "<i4"), ("light_selector", "<i4")])

SELECTOR_UNIFORM = 0     # make_uniform_light_selector (light/light_selector.art:26-44)
SELECTOR_CDF = 1         # "simple": make_cdf_light_selector over the lights' flux (light_selector.art:46-77, LoaderLight.cpp:440-476)
SELECTOR_HIERARCHY = 2   # make_hierarchy_light_selector (light_selector.art:79-110, light/light_hierarchy.art)
assert LOOKUP_DTYPE.itemsize == 16 and LEAF_DTYPE.itemsize == 96
assert MATERIAL_DTYPE.itemsize == 128 and LIGHT_DTYPE.itemsize == 128 and TEXTURE_DTYPE.itemsize == 96
assert CAMERA_DTYPE.itemsize == 56 and TECHNIQUE_DTYPE.itemsize == 20


class SceneError(RuntimeError):
    pass


@dataclass
class SceneTables:
    """Everything `igb200_set_scene` / the oracle take. All arrays are C-contiguous little-endian."""
    entities: np.ndarray            # (n_entities, 36) f32
    shape_lookups: np.ndarray       # (n_shapes,) LOOKUP_DTYPE
    shape_data: np.ndarray          # (bytes,) u8
    leaves: np.ndarray              # (n_entities,) LEAF_DTYPE
    entity_per_material: np.ndarray  # (n_materials,) i32
    materials: np.ndarray           # (n_materials,) MATERIAL_DTYPE
    infinite_lights: np.ndarray     # LIGHT_DTYPE
    finite_lights: np.ndarray       # LIGHT_DTYPE
    camera: np.ndarray              # () CAMERA_DTYPE
    technique: np.ndarray           # () TECHNIQUE_DTYPE
    bbox_min: np.ndarray            # (3,) f32
    bbox_max: np.ndarray            # (3,) f32
    film_size: tuple[int, int]
    entity_names: list[str] = field(default_factory=list)
    material_names: list[str] = field(default_factory=list)
    selector_data: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))   # light_cdf.bin / light_hierarchy.bin as f32 words
    textures: np.ndarray = field(default_factory=lambda: np.zeros(0, TEXTURE_DTYPE))
    images: list = field(default_factory=list)       # (format, array): RGBA8 -> (H, W, 4) u8, MONO8 -> (H, W) u8, RGBA32F -> (H, W, 4) f32; rows bottom-up
    image_files: list = field(default_factory=list)  # per image: (file name, linear flag) -- None for images that are no file (the baked sky)
    aux_data: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))        # 2-D cdfs of textured environment lights
    texture_names: list = field(default_factory=list)
    embedded_lights: bool = False                    # the finite lights come from embedded fix-tables (LoaderLight.h:27)
    sun_params: dict = field(default_factory=dict)    # infinite light index -> (direction before vec3_normalize, angle in degrees, radiance given?, colour as given)
    spot_angles: dict = field(default_factory=dict)   # finite light index -> (cutoff, falloff) in degrees (the descriptors hold the cosines)

    @property
    def n_entities(self) -> int:
        return int(self.entities.shape[0])

    @property
    def n_triangles(self) -> int:
        """Instanced triangle count (per entity, not per shape)."""
        total = 0
        for leaf in self.leaves:
            lk = self.shape_lookups[int(leaf["shape_id"])]
            if int(lk["type_id"]) == SHAPE_TRIMESH:
                total += int(np.frombuffer(self.shape_data, "<u4", 1, int(lk["offset"]))[0])
        return total


def spot_cosines(cutoff_deg: float, falloff_deg: float):
    """cos(rad(cutoff)), cos(rad(falloff)) of make_spot_light (light/spot.art:10-11) as f32: rad() in f32, cos in double."""
    def one(deg):
        r = F(F(deg) / F(180) * F(3.14159265359))
        return float(F(math.cos(float(r))))
    return one(cutoff_deg), one(falloff_deg)


def ellipsoid_area(global_linear, radius: float) -> float:
    """compute_ellipsoid_area (src/artic/shapes/sphere.art:21-27): Knud Thomsen's formula on the squared axis lengths, as the
    reference writes it. A per-light constant: evaluated here in double and rounded once (the C++ recogniser does the same)."""
    m = np.asarray(global_linear, np.float32).astype(np.float64)
    r = float(np.float32(radius))
    l1, l2, l3 = (float(np.dot(m[:, k] * r, m[:, k] * r)) for k in range(3))
    p = float(np.float32(1.6))
    return float(F(4 * float(np.float32(3.14159265359)) * math.pow((math.pow(l1 * l2, p / 2) + math.pow(l1 * l3, p / 2) + math.pow(l2 * l3, p / 2)) / 3, 1 / p)))


# ------------------------------------------------------------------ JSON helpers
def _load_json(path: str) -> dict:
    with open(path, "r") as fh:
        doc = json.load(fh)
    base = os.path.dirname(os.path.abspath(path))
    merged: dict = {}
    # Parser.cpp:453-463: externals first, the including file overrides/extends afterwards
    for ext in doc.get("externals", []):
        sub = _load_json(os.path.join(base, ext["filename"]))
        _merge(merged, sub)
    doc = {k: v for k, v in doc.items() if k != "externals"}
    _tag_paths(doc, base)
    _merge(merged, doc)
    return merged


def _tag_paths(doc: dict, base: str) -> None:
    for sec in ("shapes", "textures"):
        for obj in doc.get(sec, []):
            if isinstance(obj, dict) and "filename" in obj and not os.path.isabs(obj["filename"]):
                obj["filename"] = os.path.normpath(os.path.join(base, obj["filename"]))


def _merge(dst: dict, src: dict) -> None:
    for k, v in src.items():
        if k in ("bsdfs", "shapes", "entities", "lights", "textures", "media", "parameters"):
            cur = dst.setdefault(k, [])
            names = {o.get("name"): i for i, o in enumerate(cur) if isinstance(o, dict)}
            for o in v:
                if isinstance(o, dict) and o.get("name") in names:
                    cur[names[o["name"]]] = o
                else:
                    cur.append(o)
        else:
            dst[k] = v


def _vec3(v, default=None) -> np.ndarray:
    if v is None:
        return np.asarray(default, F)
    if isinstance(v, (int, float)):
        return np.asarray([v, v, v], F)
    if len(v) == 2:
        return np.asarray([v[0], v[1], 0], F)
    if len(v) != 3:
        raise SceneError("Expected vector of length 3")
    return np.asarray(v, F)


def _color(v, default) -> np.ndarray:
    if v is None:
        return np.asarray(default, F)
    if isinstance(v, str):
        m = re.fullmatch(r"\s*color\(([^()]*)\)\s*", v)   # constant PExpr colour literal, e.g. "color(0.8, 0.8, 0.8, 1.0)"
        if m:
            parts = [float(x) for x in m.group(1).split(",")]
            if len(parts) in (3, 4):
                return np.asarray(parts[:3], F)
            if len(parts) == 1:
                return np.asarray(parts * 3, F)
        raise SceneError(f"textured / expression colour '{v}' is outside the supported path (SURVEY §8f)")
    return _vec3(v)


def _look_at(eye, center, up) -> np.ndarray:
    """Parser.cpp:144-171."""
    eye, center, up = (np.asarray(x, np.float64) for x in (eye, center, up))
    f = center - eye
    fl = np.linalg.norm(f)
    f = f / fl if fl > 0 else np.array([0.0, 0.0, 1.0])
    u = up / np.linalg.norm(up)
    s = np.cross(f, u)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = s, u, f, eye
    return m


def _rot(axis: int, deg: float) -> np.ndarray:
    a = deg * math.pi / 180.0
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
    return m


def _matrix(vals) -> np.ndarray:
    n = len(vals)
    m = np.eye(4)
    if n == 9:
        m[:3, :3] = np.asarray(vals, np.float64).reshape(3, 3)
    elif n in (12, 16):
        m[: n // 4, :] = np.asarray(vals, np.float64).reshape(n // 4, 4)
    else:
        raise SceneError("Expected transform matrix of size 9, 12 or 16")
    return m


def _apply_op(t: np.ndarray, op: dict) -> np.ndarray:
    """Parser.cpp:173-232: every operation right-multiplies."""
    for name, val in op.items():
        if name == "translate":
            m = np.eye(4)
            m[:3, 3] = _vec3(val).astype(np.float64)
        elif name == "scale":
            m = np.eye(4)
            s = [val] * 3 if isinstance(val, (int, float)) else list(_vec3(val))
            m[0, 0], m[1, 1], m[2, 2] = (float(x) for x in s)
        elif name == "rotate":
            a = _vec3(val)
            m = _rot(0, float(a[0])) @ _rot(1, float(a[1])) @ _rot(2, float(a[2]))
        elif name == "qrotate":
            w, x, y, z = (float(q) for q in val)
            m = np.eye(4)
            m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        elif name == "lookat":
            origin = _vec3(val.get("origin"), (0, 0, 0))
            up = _vec3(val.get("up"), (0, 0, 1))
            if "direction" in val:
                target = _vec3(val["direction"]) + origin
            else:
                target = _vec3(val.get("target"), (0, 1, 0))
            m = _look_at(origin, target, up)
        elif name == "matrix":
            m = _matrix(val)
        else:
            raise SceneError(f"Transform property got unknown entry type '{name}'")
        t = t @ m
    return t


def parse_transform(v) -> np.ndarray:
    """Parser.cpp:285-320: array of numbers (row-major 3x3 / 3x4 / 4x4), list of operations, or one operation object."""
    if v is None:
        return np.eye(4)
    if isinstance(v, dict):
        return _apply_op(np.eye(4), v)
    if isinstance(v, list) and v and isinstance(v[0], dict):
        t = np.eye(4)
        for op in v:
            t = _apply_op(t, op)
        return t
    return _matrix(v)


def _bbox_transformed(lo, hi, m) -> tuple[np.ndarray, np.ndarray]:
    c = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], np.float64)
    c = c @ m[:3, :3].T + m[:3, 3]
    return c.min(axis=0).astype(F), c.max(axis=0).astype(F)


# ------------------------------------------------------------------ shapes
def _build_trimesh(obj: dict) -> TriMesh:
    """TriMeshProvider.cpp:17-104,478-548."""
    t = obj["type"].lower()
    g = obj.get
    if t == "triangle":
        m = meshes.make_triangle(_vec3(g("p0"), (0, 0, 0)), _vec3(g("p1"), (1, 0, 0)), _vec3(g("p2"), (0, 1, 0)))
    elif t == "rectangle":
        if "p0" not in obj:
            w, h = float(g("width", 2.0)), float(g("height", 2.0))
            origin = _vec3(g("origin"), (-w / 2, -h / 2, 0))
            m = meshes.make_plane(origin, (w, 0, 0), (0, h, 0))
        else:
            m = meshes.make_rectangle(_vec3(g("p0"), (-1, -1, 0)), _vec3(g("p1"), (1, -1, 0)),
                                      _vec3(g("p2"), (1, 1, 0)), _vec3(g("p3"), (-1, 1, 0)))
    elif t in ("cube", "box"):
        w, h, d = float(g("width", 2.0)), float(g("height", 2.0)), float(g("depth", 2.0))
        origin = _vec3(g("origin"), (-w / 2, -h / 2, -d / 2))
        m = meshes.make_box(origin, (w, 0, 0), (0, h, 0), (0, 0, d))
    elif t == "icosphere":
        m = meshes.make_ico_sphere(_vec3(g("center"), (0, 0, 0)), float(g("radius", 1.0)), int(g("subdivisions", 4)))
    elif t == "uvsphere":
        m = meshes.make_uv_sphere(_vec3(g("center"), (0, 0, 0)), float(g("radius", 1.0)), int(g("stacks", 32)), int(g("slices", 16)))
    elif t == "cylinder":
        if "radius" in obj:
            br = tr = float(g("radius", 1.0))
        else:
            br = float(g("bottom_radius", 1.0))
            tr = float(g("top_radius", br))
        m = meshes.make_cylinder(_vec3(g("p0"), (0, 0, 0)), br, _vec3(g("p1"), (0, 0, 1)), tr, int(g("sections", 32)), bool(g("filled", True)))
    elif t == "cone":
        m = meshes.make_cone(_vec3(g("p0"), (0, 0, 0)), float(g("radius", 1.0)), _vec3(g("p1"), (0, 0, 1)), int(g("sections", 32)), bool(g("filled", True)))
    elif t == "disk":
        m = meshes.make_disk(_vec3(g("origin"), (0, 0, 0)), _vec3(g("normal"), (0, 0, 1)), float(g("radius", 1.0)), int(g("sections", 32)))
    elif t in ("obj", "ply", "external"):
        si = g("shape_index", -1)
        m = meshes.load_external(obj["filename"], None if si is None or si < 0 else int(si))
    elif t == "inline":
        idx = np.asarray(g("indices"), np.int64).reshape(-1, 3)
        m = TriMesh()
        m.vertices = np.asarray(g("vertices"), F).reshape(-1, 3)
        m.indices = np.concatenate([idx, np.zeros((len(idx), 1), np.int64)], axis=1).astype(np.uint32)
        if "normals" in obj:
            m.normals = np.asarray(g("normals"), F).reshape(-1, 3)
        else:
            m.compute_vertex_normals()
        if "texcoords" in obj:
            m.texcoords = np.asarray(g("texcoords"), F).reshape(-1, 2)
        else:
            m.make_texcoords_normalized()
    else:
        raise SceneError(f"Can not load shape type '{t}'")
    if m.face_count == 0 or len(m.vertices) == 0:
        raise SceneError(f"Shape '{obj.get('name')}': no geometry generated")
    if g("flip_normals", False):
        m.flip_normals()
    if g("face_normals", False):
        m.setup_face_normals_as_vertex_normals()
    elif g("smooth_normals", False):
        m.compute_vertex_normals()
    if g("generic_uv", False):
        m.make_texcoords_normalized()
    m.transform(parse_transform(g("transform")))
    # handleModification, TriMeshProvider.cpp:431-476
    sub = int(g("subdivision", 0))
    if float(g("refinement", 0)) > 0 or g("displacement"):
        raise SceneError("refinement / displacement are outside the supported path")
    if sub > 0:
        for _ in range(sub):
            m.subdivide()
        if g("smooth_normals", False):
            m.compute_vertex_normals()
        else:
            m.setup_face_normals_as_vertex_normals()
    return m


def _serialize_trimesh(m: TriMesh, lo: np.ndarray, hi: np.ndarray) -> bytes:
    """TriMeshProvider.cpp:575-596."""
    nf, nv, nn, nt = m.face_count, len(m.vertices), len(m.normals), len(m.texcoords)
    head = np.asarray([nf, nv, nn, nt], "<u4").tobytes()
    box = np.asarray([lo[0], lo[1], lo[2], 0, hi[0], hi[1], hi[2], 0], "<f4").tobytes()
    v4 = np.zeros((nv, 4), "<f4")
    v4[:, :3] = m.vertices
    n4 = np.zeros((nn, 4), "<f4")
    n4[:, :3] = m.normals
    return head + box + v4.tobytes() + n4.tobytes() + m.indices.astype("<u4").tobytes() + m.texcoords.astype("<f4").tobytes()


# ------------------------------------------------------------------ images and textures
def read_png(path: str) -> np.ndarray:
    """Minimal PNG decoder (8-bit grey / grey+alpha / RGB / RGBA, non-interlaced) -> (H, W, C) u8, rows top-down as in the file."""
    import struct
    import zlib
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise SceneError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or interlace != 0 or ctype not in (0, 2, 4, 6):
        raise SceneError(f"{path}: only 8-bit non-interlaced grey / RGB / RGBA PNG files are supported")
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(h, 1 + w * ch)
    out = np.zeros((h, w * ch), np.uint8)
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:      # Sub: prefix sums per channel
            cur = line.reshape(w, ch).cumsum(axis=0).reshape(-1) & 255
        else:              # Average / Paeth depend on the reconstructed left neighbour: pixel by pixel
            cur = np.zeros(w * ch, np.int32)
            for i in range(w * ch):
                a = cur[i - ch] if i >= ch else 0
                b = prev[i]
                if ft == 3:
                    cur[i] = (line[i] + ((a + b) >> 1)) & 255
                else:
                    c = prev[i - ch] if i >= ch else 0
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    cur[i] = (line[i] + (a if pa <= pb and pa <= pc else (b if pb <= pc else c))) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, ch)


def srgb_byte_to_linear_byte() -> np.ndarray:
    """byte_color_to_linear for all 256 values (src/runtime/Image.cpp:40-51): floor(srgb_invgamma(c / 255) * 255) in f32."""
    c = (np.arange(256, dtype=F) / F(255)).astype(F)
    lin = np.where(c <= F(0.04045), c / F(12.92), np.power(((c + F(0.055)) / F(1.055)).astype(F), F(2.4)).astype(F)).astype(F)
    return np.minimum(255, np.floor(lin * F(255))).astype(np.uint8)


def load_image_for_device(path: str, linear_hint: bool = False):
    """An image as the reference's device holds it (igb200_image): (format, array), rows bottom-up.
    8-bit files are `packed` (Image::isPacked, Image.cpp:358-365) and mapped to linear bytes unless the texture says `linear`
    (Image::loadAsPacked, Image.cpp:714-810 -- NB its `linear` argument means "already linear"); .npy / .npz hold float RGB(A)
    fixtures made from the reference's EXR files by tools/make_golden.py (Image::load, Image.cpp:500-712)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".png":
        px = read_png(path)[::-1]                               # stbi_set_flip_vertically_on_load(1)
        lut = np.arange(256, dtype=np.uint8) if linear_hint else srgb_byte_to_linear_byte()
        h, w, ch = px.shape
        if ch == 1:
            return IMAGE_MONO8, np.ascontiguousarray(lut[px[:, :, 0]])
        if ch == 2:                                             # grey + alpha: stb re-reads it as RGBA (Image.cpp:730-734)
            px = np.concatenate([px[:, :, :1]] * 3 + [px[:, :, 1:]], axis=2)
        out = np.full((h, w, 4), 255, np.uint8)
        out[:, :, :3] = lut[px[:, :, :3]]
        if px.shape[2] == 4:
            out[:, :, 3] = px[:, :, 3]
        return IMAGE_RGBA8, np.ascontiguousarray(out)
    if ext in (".npy", ".npz"):
        a = np.load(path)
        if ext == ".npz":
            a = a[a.files[0]]
        a = np.asarray(a, F)
        if a.ndim != 3 or a.shape[2] not in (3, 4):
            raise SceneError(f"{path}: expected an (H, W, 3|4) array, rows top-down as in the image file")
        out = np.ones((a.shape[0], a.shape[1], 4), F)
        out[:, :, :a.shape[2]] = a
        return IMAGE_RGBA32F, np.ascontiguousarray(out[::-1])   # Image::flipY
    raise SceneError(f"image file type '{ext}' is outside the supported path (PNG, or .npy / .npz fixtures of float images)")


def transform_2d(v) -> np.ndarray:
    """LoaderUtils::inlineTransformAs2d (LoaderUtils.cpp:40-46): rows 0 and 1 of [linear 2x2 | translation xy]."""
    t = parse_transform(v)
    # the generator streams the matrix into shader text (LoaderUtils::inlineMatrix / inlineVector: operator<< with the default precision), so
    # the device only ever sees six significant digits of every element
    return np.array([float("%g" % float(F(x))) for x in (t[0, 0], t[0, 1], t[0, 3], t[1, 0], t[1, 1], t[1, 3])], F)


def image_pixels_f32(fmt: int, arr: np.ndarray) -> np.ndarray:
    """(H, W, 3) f32 pixel values of a device image (driver/image.art:9-34)."""
    if fmt == IMAGE_RGBA32F:
        return arr[:, :, :3].astype(F)
    if fmt == IMAGE_MONO8:
        g = (arr.astype(F) / F(255)).astype(F)
        return np.stack([g, g, g], axis=2)
    return (arr[:, :, :3].astype(F) / F(255)).astype(F)


def _border_np(mode: int, x: np.ndarray, w: int) -> np.ndarray:
    if mode == 1:
        return np.clip(x, 0, w - 1)
    if mode == 2:
        t = np.where(x < 0, -1 - x, x)
        i = t // w
        k = t - i * w
        return np.where((i & 1) == 0, w - 1 - k, k)
    t = np.fmod(x, w)            # C remainder (sign of the dividend), as Artic's % on i32
    return np.where(t < 0, t + w, t)


def eval_texture_np(tex, images, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Host-side texture evaluation on arrays of uv (f32), for baking (what the reference does with ig_bake_shader,
    entrypoints/bake.art:1-27). Same formulas as the device (texture/checkerboard.art, texture/image.art). Returns (..., 3) f32."""
    u, v = np.asarray(u, F), np.asarray(v, F)
    m = tex["transform"].astype(F)
    fma = lambda a, b, c: (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(F)   # one rounding, as fmaf
    u2 = fma(u, m[0], fma(v, m[1], np.full_like(u, m[2] * F(1))))
    v2 = fma(u, m[3], fma(v, m[4], np.full_like(u, m[5] * F(1))))
    if int(tex["type"]) == TEX_CHECKERBOARD:
        p = tex["p"].astype(F)
        wrap2 = lambda a: (a - F(2) * np.floor(a / F(2))).astype(F)
        px = (wrap2((u2 * p[0]).astype(F)).astype(np.int32) % 2) == 0
        py = (wrap2((v2 * p[1]).astype(F)).astype(np.int32) % 2) == 0
        return np.where((px != py)[..., None], p[2:5], p[5:8]).astype(F)
    fmt, arr = images[int(tex["image"])]
    pix = image_pixels_f32(fmt, arr)
    h, w = pix.shape[:2]
    flt, bu, bv = int(tex["filter"]), int(tex["border_u"]), int(tex["border_v"])
    px = lambda x, y: pix[_border_np(bv, y, h), _border_np(bu, x, w)]
    if flt == 0:
        return px(np.floor(u2 * F(w)).astype(np.int64), np.floor(v2 * F(h)).astype(np.int64))
    uu, vv = (u2 * F(w) - F(0.5)).astype(F), (v2 * F(h) - F(0.5)).astype(F)
    ix, iy = np.floor(uu).astype(np.int64), np.floor(vv).astype(np.int64)
    fx, fy = (uu - np.floor(uu)).astype(F)[..., None], (vv - np.floor(vv)).astype(F)[..., None]
    lerp = lambda a, b, t: ((F(1) - t) * a + t * b).astype(F)
    if flt == 1:
        return lerp(lerp(px(ix, iy), px(ix + 1, iy), fx), lerp(px(ix, iy + 1), px(ix + 1, iy + 1), fx), fy)
    w0 = lambda a: (a * (a * (-a + F(3)) - F(3)) + F(1)) / F(6)
    w1 = lambda a: (a * a * (F(3) * a - F(6)) + F(4)) / F(6)
    w2 = lambda a: (a * (a * (F(-3) * a + F(3)) + F(3)) + F(1)) / F(6)
    w3 = lambda a: (a * a * a) / F(6)
    g0, g1 = (lambda a: w0(a) + w1(a)), (lambda a: w2(a) + w3(a))
    h0, h1 = (lambda a: (w1(a) / g0(a)) - F(1)), (lambda a: (w3(a) / g1(a)) + F(1))
    fx1, fy1 = fx[..., 0], fy[..., 0]
    ix0 = np.floor(ix.astype(F) + h0(fx1) + F(0.5)).astype(np.int64); iy0 = np.floor(iy.astype(F) + h0(fy1) + F(0.5)).astype(np.int64)
    ix1 = np.floor(ix.astype(F) + h1(fx1) + F(0.5)).astype(np.int64); iy1 = np.floor(iy.astype(F) + h1(fy1) + F(0.5)).astype(np.int64)
    return ((px(ix0, iy0) * (g0(fx) * g0(fy)) + px(ix1, iy0) * (g1(fx) * g0(fy))) + (px(ix0, iy1) * (g0(fx) * g1(fy)) + px(ix1, iy1) * (g1(fx) * g1(fy)))).astype(F)


def bake_texture(tex, images, max_w: int = 1024, max_h: int = 512) -> np.ndarray:
    """ShadingTree::bakeTexture with TextureBakeOptions{0, 0, 1024, 512} (ShadingTree.cpp:580-630, EnvironmentLight.cpp:52): the texture on a
    grid uv = (x / (w - 1), y / (h - 1)), resolution = the image's, capped. Returns (h, w, 3) f32."""
    if int(tex["type"]) == TEX_IMAGE:
        fmt, arr = images[int(tex["image"])]
        h, w = arr.shape[:2]
    else:
        w = h = 1
    w, h = max(1, min(max_w, w)), max(1, min(max_h, h))
    xs = (np.arange(w, dtype=F) / F(max(w - 1, 1))).astype(F) if w > 1 else np.zeros(1, F)
    ys = (np.arange(h, dtype=F) / F(max(h - 1, 1))).astype(F) if h > 1 else np.zeros(1, F)
    uu, vv = np.meshgrid(xs, ys)
    return eval_texture_np(tex, images, uu, vv)


SKY_KEYS = ("ground", "turbidity", "direction", "sun_direction", "elevation", "azimuth", "year", "month", "day", "hour", "minute", "seconds",
            "latitude", "longitude", "timezone")
SKY_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes", "textures", "sky")


def sky_key(lj: dict) -> str:
    import hashlib
    return hashlib.sha1(json.dumps({k: lj[k] for k in SKY_KEYS if k in lj}, sort_keys=True).encode()).hexdigest()[:12]


def sky_image(lj: dict) -> np.ndarray:
    """The sky model image of a `sky` light as the device holds it: (256, 512, 4) f32, rows bottom-up. The model itself (Hosek-Wilkie
    coefficient tables, sun position from date and place: src/runtime/skysun/) is loader-side data production outside the hot path;
    its output for a given parameter set is a fixture under scenes/textures/sky/ made by tools/make_sky.py from the reference's sources."""
    path = os.path.join(SKY_DIR, f"sky_{sky_key(lj)}.npz")
    if not os.path.exists(path):
        raise SceneError(f"no baked sky image for these parameters ({path}); run tools/make_sky.py with the reference tree present")
    a = np.load(path)["rgb"].astype(F)       # file order: row 0 = zenith (SkyModel.cpp:30-52)
    out = np.ones((a.shape[0], a.shape[1], 4), F)
    out[:, :, :3] = a
    return np.ascontiguousarray(out[::-1])   # Image::load flips (Image.cpp:708)


def cdf2d_for_image(rgb: np.ndarray, premultiply_sin: bool = True, compensate: bool = True) -> np.ndarray:
    """CDF::computeForImage (src/runtime/CDF.cpp:46-150): [marginal (H) | H conditionals (W)], each without the leading 0.
    `rgb`: (H, W, 3) f32 in memory order (rows as the device sees them)."""
    rgb = np.asarray(rgb, F)
    h, w, _ = rgb.shape
    resp = lambda a: ((np.maximum(a[..., 0], F(0)) + np.maximum(a[..., 1], F(0)) + np.maximum(a[..., 2], F(0))) / F(3)).astype(F)
    defect = F(0)
    if compensate:   # computeMISDefect: sequential f32 sum of response / width over all pixels, then / height
        r = resp(rgb).reshape(-1)
        d = F(np.cumsum((r / F(w)).astype(F), dtype=F)[-1] / F(h))
        defect = F(0) if abs(float(r.min()) - float(d)) < 1e-4 else d
    rr = resp((rgb - defect).astype(F))
    cond = np.cumsum(rr, axis=1, dtype=F)                      # np.add.accumulate is sequential in the array's dtype
    total = cond[:, -1].copy()
    sin_row = np.sin((F(PI) * (np.arange(h, dtype=F) + F(0.5)) / F(h)).astype(F)).astype(F)
    marg = (total * sin_row).astype(F) if premultiply_sin else total.copy()
    ok = total > F(1e-5)
    cond[ok] = (cond[ok] * (F(1) / total[ok])[:, None]).astype(F)
    if (~ok).any():
        ramp = (np.arange(1, w + 1, dtype=F) * F(F(1) / F(w))).astype(F)
        cond[~ok] = ramp
    cond[:, -1] = 1
    marg = np.cumsum(marg, dtype=F)
    if marg[-1] > F(1e-5):
        marg = (marg * F(F(1) / marg[-1])).astype(F)
    else:
        marg = (np.arange(1, h + 1, dtype=F) * F(F(1) / F(h))).astype(F)
    marg[-1] = 1
    return np.concatenate([marg, cond.reshape(-1)]).astype(F)


# ------------------------------------------------------------------ loader
def light_direction(lj) -> np.ndarray:
    """LoaderUtils::getDirection (src/runtime/loader/LoaderUtils.cpp:89-106, skysun/ElevationAzimuth.h:15-30): `direction` (or
    `sun_direction`) goes through elevation / azimuth and back (y up), else `elevation` / `azimuth` are used as given (radians)."""
    if "direction" in lj or "sun_direction" in lj:
        d = np.asarray(lj.get("direction", lj.get("sun_direction")), np.float64)
        d = (d / np.linalg.norm(d)).astype(F)
        theta = F(math.acos(float(d[1])))
        phi = F(math.atan2(-float(d[0]), -float(d[2])))
        elev, azim = F(F(math.pi / 2) - theta), (F(phi + F(2 * math.pi)) if phi < 0 else phi)
    elif "elevation" in lj or "azimuth" in lj:
        elev, azim = F(lj.get("elevation", 0)), F(lj.get("azimuth", 0))
    else:
        raise SceneError("sun position from time and location is outside the supported path")
    se, ce, sa, ca = (F(f(float(x))) for f, x in ((math.sin, elev), (math.cos, elev), (math.sin, azim), (math.cos, azim)))
    return np.array([-ce * sa, se, -ce * ca], F)


def normalize_f32(v) -> np.ndarray:
    """vec3_normalize as the C++ recogniser evaluates it: the length in double, rounded once, then three f32 divisions."""
    v = np.asarray(v, F)
    return (v / F(math.sqrt(float(v[0]) ** 2 + float(v[1]) ** 2 + float(v[2]) ** 2))).astype(F)


def light_cdf(flux) -> np.ndarray:
    """`light_cdf.bin` of the "simple" selector: CDF::computeForArray (src/runtime/CDF.cpp:14-44) over the lights' flux. The leading
    0 is not stored: [x1, ..., x(n-1), 1] (core/cdf.art:70-73)."""
    v = np.asarray(flux, F)
    cdf = np.zeros(len(v), F)
    acc = F(0)
    for i, x in enumerate(v):
        acc = F(acc + x) if i else F(x)
        cdf[i] = acc
    total = cdf[-1]
    if total > F(1e-5):
        cdf = (cdf * F(F(1) / total)).astype(F)
    else:
        n = F(F(1) / F(len(v)))
        for x in range(1, len(v)):
            cdf[x - 1] = F(F(x) * n)
    cdf[-1] = 1
    return cdf


def light_hierarchy(lights) -> np.ndarray:
    """`light_hierarchy.bin` (src/runtime/light/LightHierarchy.cpp:47-129): a binary tree over the finite lights' positions, built
    by inserting them one by one (src/runtime/container/PointBvh.inl:23-96), every node carrying a position, a direction and a
    flux (negative = no direction). Returns the file's f32 words: codes[round_up(n, 4)] (u32: the left/right turns from the root
    to the light, LSB first) followed by 8 words per node {pos xyz, flux, dir xyz, id} (id >= 0: light; < 0: -(left child + 1)).

    `lights`: (position, direction or None, flux) per finite light, in table order."""
    inf = float("inf")
    nodes = []   # dict(index, lo, hi, mid, axis): axis < 0 = leaf holding light `index`

    def leaf(lo, hi):
        return dict(index=0, lo=np.array(lo, F), hi=np.array(hi, F), mid=F(0), axis=-1)

    for li, (pos, _, _) in enumerate(lights):
        p = np.asarray(pos, F)
        if not nodes:
            nodes.append(leaf(p, p))
            continue
        # descend to the leaf the point falls into, growing every box on the way (getForPointExtend)
        k = 0
        while True:
            nd = nodes[k]
            nd["lo"], nd["hi"] = np.minimum(nd["lo"], p), np.maximum(nd["hi"], p)
            if nd["axis"] < 0:
                break
            ctr = F((nd["hi"][nd["axis"]] + nd["lo"][nd["axis"]]) / F(2))
            k = nd["index"] if p[nd["axis"]] < ctr else nd["index"] + 1
        old = nd["index"]
        diam = (nd["hi"] - nd["lo"]).astype(F)
        axis = int(np.argmax(diam))                      # Eigen maxCoeff: first maximum
        mid = F(diam[axis] / F(2))                       # half the extent (the reference compares the coordinate with this value)
        left_idx = len(nodes)
        nd["index"], nd["axis"], nd["mid"] = left_idx, axis, mid
        off = F((nd["hi"][axis] - nd["lo"][axis]) * F(0.5))
        lhi = nd["hi"].copy(); lhi[axis] = F(lhi[axis] - off)
        rlo = nd["lo"].copy(); rlo[axis] = F(rlo[axis] + off)
        left, right = leaf(nd["lo"], lhi), leaf(rlo, nd["hi"])
        if p[axis] < mid:
            left["index"], right["index"] = li, old
        else:
            left["index"], right["index"] = old, li
        nodes += [left, right]

    n = len(lights)
    entries = np.zeros((len(nodes), 8), F)
    codes = np.zeros((n + 3) // 4 * 4, np.uint32)

    def populate(k, code, depth):   # populateInnerNodes; returns (dir, signed flux)
        nd = nodes[k]
        e = entries[k]
        if nd["axis"] < 0:
            pos, d, flux = lights[nd["index"]]
            e[0:3] = pos
            e[3] = flux if d is not None else -flux
            e[4:7] = d if d is not None else (0, 0, 1)
            e[7:8].view(np.int32)[0] = nd["index"]
            codes[nd["index"]] = code
        else:
            ld, lf = populate(nd["index"], code, depth + 1)
            rd, rf = populate(nd["index"] + 1, code | (1 << depth), depth + 1)
            e[0:3] = ((nd["hi"] + nd["lo"]) / F(2)).astype(F)
            e[7:8].view(np.int32)[0] = -(nd["index"] + 1)
            if lf < 0 and rf < 0:
                e[4:7], e[3] = (0, 0, 1), F(lf + rf)
            elif lf < 0:
                e[4:7], e[3] = (0, 0, 1), F(-(F(-lf) + rf))
            elif rf < 0:
                e[4:7], e[3] = (0, 0, 1), F(-(lf - rf))
            else:
                sm = (ld + rd).astype(F)
                nrm = F(np.sqrt(np.sum(sm * sm, dtype=F)))
                e[4:7], e[3] = (sm / nrm if nrm > 0 else sm), F(lf + rf)
        return e[4:7].copy(), F(e[3])

    populate(0, 0, 0)
    return np.concatenate([codes.view(F), entries.reshape(-1)]).astype(F)


def load_scene(path, width: int | None = None, height: int | None = None,
               max_depth: int | None = None, base_dir: str | None = None) -> SceneTables:
    """`path` is a scene file name or an already parsed scene dict (the reference's `loadFromString`)."""
    if isinstance(path, dict):
        doc = json.loads(json.dumps(path))
        base = base_dir or os.getcwd()
        merged: dict = {}
        for ext in doc.get("externals", []):
            _merge(merged, _load_json(os.path.join(base, ext["filename"])))
        doc = {k: v for k, v in doc.items() if k != "externals"}
        _tag_paths(doc, base)
        _merge(merged, doc)
        doc = merged
    else:
        doc = _load_json(path)
    film = doc.get("film", {})
    fsize = film.get("size", [800, 600])
    fw = int(width if width is not None else fsize[0])
    fh = int(height if height is not None else fsize[1])

    tech = doc.get("technique", {"type": "path"})
    if tech.get("type", "path") != "path":
        raise SceneError(f"technique '{tech.get('type')}' is outside the supported path (only 'path')")
    if tech.get("aov_mis", False):
        raise SceneError("aov_mis (advanced shadow handling) is outside the supported path")
    sel = str(tech.get("light_selector", "") or "uniform").lower()
    technique = np.zeros((), TECHNIQUE_DTYPE)
    technique["max_depth"] = int(max_depth if max_depth is not None else tech.get("max_depth", 64))
    technique["min_depth"] = int(tech.get("min_depth", 2))
    technique["clamp"] = float(tech.get("clamp", 0.0))
    technique["nee"] = 1 if tech.get("nee", True) else 0

    bsdfs = {b["name"]: b for b in doc.get("bsdfs", [])}

    # ---- textures (LoaderTexture.cpp:60-70, pattern/{Image,CheckerBoard}Pattern.cpp): declared lazily, in order of first use
    tex_json = {t["name"]: t for t in doc.get("textures", [])}
    tex_ids: dict[str, int] = {}
    tex_recs, images, image_ids, aux_words = [], [], {}, []
    image_files: list = []

    def texture_id(name: str) -> int:
        if name in tex_ids:
            return tex_ids[name]
        tj = tex_json[name]
        tt = tj.get("type", "").lower()
        rec = np.zeros((), TEXTURE_DTYPE)
        rec["transform"] = transform_2d(tj.get("transform"))
        if tt in ("image", "bitmap"):
            fname = tj["filename"]
            key = (fname, bool(tj.get("linear", False)))
            if key not in image_ids:
                image_ids[key] = len(images)
                images.append(load_image_for_device(fname, key[1]))
                image_files.append((os.path.abspath(fname), key[1]))
            rec["type"], rec["image"] = TEX_IMAGE, image_ids[key]
            rec["filter"] = FILTERS.get(str(tj.get("filter_type", "bicubic")), FILTER_BICUBIC)
            if "wrap_mode_u" in tj:
                rec["border_u"] = BORDERS.get(str(tj.get("wrap_mode_u", "repeat")), 0)
                rec["border_v"] = BORDERS.get(str(tj.get("wrap_mode_v", "repeat")), 0)
            else:
                rec["border_u"] = rec["border_v"] = BORDERS.get(str(tj.get("wrap_mode", "repeat")), 0)
        elif tt == "checkerboard":
            rec["type"] = TEX_CHECKERBOARD
            rec["p"][0], rec["p"][1] = float(tj.get("scale_x", 2.0)), float(tj.get("scale_y", 2.0))
            rec["p"][2:5] = _color(tj.get("color0"), (0, 0, 0))
            rec["p"][5:8] = _color(tj.get("color1"), (1, 1, 1))
        else:
            raise SceneError(f"texture type '{tt}' is outside the supported path (image / bitmap, checkerboard)")
        tex_ids[name] = len(tex_recs)
        tex_recs.append(rec)
        return tex_ids[name]

    def color_or_tex(v, default):
        """A colour property: (rgb, texture id or -1). A string naming a texture is the PExpr `name` = tex_name(uv)."""
        if isinstance(v, str) and v.strip() in tex_json:
            return np.zeros(3, F), texture_id(v.strip())
        return _color(v, default), -1
    shapes_json = {s["name"]: s for s in doc.get("shapes", [])}
    entities_json = doc.get("entities", [])
    lights_json = doc.get("lights", [])

    # LoaderLight.cpp:263-277: entities referenced by an area light are emissive
    emissive = {l["entity"]: l for l in lights_json if l.get("type") == "area" and l.get("entity")}

    # ---- shapes actually used, in file order (LoaderShape.cpp prepare: only referenced shapes are loaded)
    used = []
    for e in entities_json:
        s = e.get("shape")
        if s in shapes_json and s not in used:
            used.append(s)
    shape_ids, shape_info = {}, []
    lookups, blob = [], bytearray()
    for name in used:
        sj = shapes_json[name]
        off = len(blob)
        if sj["type"].lower() == "sphere":
            origin = _vec3(sj.get("center"), (0, 0, 0))
            radius = F(sj.get("radius", 1.0))
            if radius <= 0:
                raise SceneError(f"Shape '{name}': invalid radius")
            lo, hi = (origin - radius - F(1e-5)).astype(F), (origin + radius + F(1e-5)).astype(F)
            blob += np.asarray([origin[0], origin[1], origin[2], radius], "<f4").tobytes()
            lookups.append((SHAPE_SPHERE, 0, off))
            shape_info.append(dict(type=SHAPE_SPHERE, lo=lo, hi=hi, mesh=None, plane=None, origin=origin, radius=radius))
        else:
            mesh = _build_trimesh(sj)
            lo = (mesh.vertices.min(axis=0) - F(1e-5)).astype(F)
            hi = (mesh.vertices.max(axis=0) + F(1e-5)).astype(F)
            blob += _serialize_trimesh(mesh, lo, hi)
            lookups.append((SHAPE_TRIMESH, 0, off))
            shape_info.append(dict(type=SHAPE_TRIMESH, lo=lo, hi=hi, mesh=mesh, plane=mesh.get_as_plane()))
        while len(blob) % 16:
            blob += b"\0"
        shape_ids[name] = len(shape_ids)

    # ---- material grouping (LoaderEntity.cpp:42-97): one material per distinct bsdf, one per emissive entity
    mat_keys: list[tuple] = []
    groups: list[list[dict]] = []
    for e in entities_json:
        bname = e.get("bsdf", "")
        if not bname or bname not in bsdfs:
            raise SceneError(f"Entity {e.get('name')} has no/unknown bsdf '{bname}'")
        if e.get("inner_medium") or e.get("outer_medium"):
            raise SceneError("participating media are outside the supported path")
        if e["name"] in emissive:
            mat_keys.append((bname, e["name"]))
            groups.append([e])
        else:
            key = (bname, None)
            if key in mat_keys:
                groups[mat_keys.index(key)].append(e)
            else:
                mat_keys.append(key)
                groups.append([e])

    n_ent = sum(len(g) for g in groups)
    ent_table = np.zeros((n_ent, 36), F)
    leaves = np.zeros(n_ent, LEAF_DTYPE)
    names, ent_info = [], {}
    bb_lo = np.full(3, np.inf, F)
    bb_hi = np.full(3, -np.inf, F)
    eid = 0
    for mat_id, grp in enumerate(groups):
        for e in grp:
            sname = e.get("shape", "")
            if sname not in shape_ids:
                raise SceneError(f"Entity {e.get('name')} has unknown shape '{sname}'")
            sid = shape_ids[sname]
            flags = 0
            for bit, key in ((1, "camera_visible"), (2, "light_visible"), (4, "bounce_visible"), (8, "shadow_visible")):
                if e.get(key, True):
                    flags |= bit
            t = parse_transform(e.get("transform"))
            t[3, :] = (0, 0, 0, 1)
            inv = np.linalg.inv(t)
            lo, hi = _bbox_transformed(shape_info[sid]["lo"], shape_info[sid]["hi"], t)
            bb_lo, bb_hi = np.minimum(bb_lo, lo), np.maximum(bb_hi, hi)
            to_local = inv[:3, :4].astype(F)
            to_global = t[:3, :4].astype(F)
            to_normal = np.linalg.inv(t[:3, :3]).T.astype(F)
            row = ent_table[eid]
            row[0:12] = to_local.T.reshape(-1)      # column-major
            row[12:24] = to_global.T.reshape(-1)
            row[24:33] = to_normal.T.reshape(-1)
            row[33:36].view(np.uint32)[:] = (sid, mat_id, 0)
            lf = leaves[eid]
            lf["min"], lf["max"] = lo, hi
            lf["entity_id"], lf["shape_id"] = eid, sid
            lf["local"] = to_local.T.reshape(-1)
            lf["flags"], lf["mat_id"] = flags, mat_id
            names.append(e["name"])
            ent_info[e["name"]] = dict(id=eid, transform=t, shape_id=sid, mat_id=mat_id)
            eid += 1

    # ---- lights (LoaderLight.cpp:263-300: infinite and finite lights have separate id spaces)
    inf_l, fin_l = [], []
    fin_sel = []   # per finite light: (position, direction or None, flux) -- Light::position/direction/computeFlux, for the light selectors
    spot_angles: dict = {}
    point_raw: dict = {}
    sun_params: dict = {}
    fin_of_entity: dict[str, int] = {}
    for lj in lights_json:
        lt = lj.get("type", "").lower()
        rec = np.zeros((), LIGHT_DTYPE)
        rec["entity_id"] = -1
        if lt in ("env", "constant", "uniform", "envmap"):
            rad = lj.get("radiance", [1, 1, 1])
            scale = _color(lj.get("scale"), (1, 1, 1))
            rgb, tid = color_or_tex(rad, (1, 1, 1))
            if tid < 0:
                rec["type"] = LIGHT_ENV_CONST
                rec["p"][0:3] = (scale * rgb).astype(F)   # color_mul(scale, tex)
            else:
                # EnvironmentLight.cpp:45-110: a textured environment; with a cdf (default "conditional", MIS compensation on) the
                # light is make_environment_light_textured over the 2-D cdf of the baked radiance, else make_environment_light
                if "transform" in lj:
                    raise SceneError("transformed environment lights are outside the supported path")
                method = str(lj.get("cdf", "conditional")).lower()
                if method not in ("none", "conditional", ""):
                    raise SceneError(f"environment cdf method '{method}' is outside the supported path (conditional, none)")
                rec["p"][0:3] = scale
                rec["p"][3:12] = (1, 0, 0, 0, 1, 0, 0, 0, 1)
                rec["p"][12:13].view(np.int32)[0] = tid
                baked = bake_texture(tex_recs[tid], images)
                if method == "none" or baked.shape[0] <= 1 or baked.shape[1] <= 1:
                    rec["type"] = LIGHT_ENV_TEX
                else:
                    rec["type"] = LIGHT_ENV_TEXTURED
                    cdf = cdf2d_for_image(baked, True, bool(lj.get("compensate", True)))
                    rec["p"][13:16].view(np.int32)[:] = (sum(len(a) for a in aux_words), baked.shape[1], baked.shape[0])
                    aux_words.append(cdf)
            inf_l.append(rec)
        elif lt == "sky":
            # SkyLight.cpp:9-80: the Hosek-Wilkie model baked into a 512 x 256 image (skysun/SkyModel.cpp), then
            # make_environment_light_textured(bilinear, repeat) over its 2-D cdf (premultiplied by sin, no compensation)
            if "transform" in lj:
                raise SceneError("transformed sky lights are outside the supported path")
            sky = sky_image(lj)
            images.append((IMAGE_RGBA32F, sky))
            image_files.append(None)
            trec = np.zeros((), TEXTURE_DTYPE)
            trec["type"], trec["image"], trec["filter"] = TEX_IMAGE, len(images) - 1, 1
            trec["transform"] = (1, 0, 0, 0, 1, 0)
            tex_recs.append(trec)
            tid = len(tex_recs) - 1
            tex_ids["_sky_" + str(lj.get("name", len(inf_l)))] = tid
            rec["type"] = LIGHT_ENV_TEXTURED
            rec["p"][0:3] = _color(lj.get("scale"), (1, 1, 1))
            rec["p"][3:12] = (1, 0, 0, 0, 1, 0, 0, 0, 1)
            rec["p"][12:13].view(np.int32)[0] = tid
            cdf = cdf2d_for_image(sky[:, :, :3], True, False)
            rec["p"][13:16].view(np.int32)[:] = (sum(len(a) for a in aux_words), sky.shape[1], sky.shape[0])
            aux_words.append(cdf)
            inf_l.append(rec)
        elif lt in ("sun", "directional", "direction", "distant"):
            # LoaderUtils::getDirection (LoaderUtils.cpp:89-106): direction -> elevation / azimuth -> direction (y up), then
            # vec3_normalize in the script (SunLight.cpp:44, DirectionalLight.cpp:36)
            d = light_direction(lj)
            rec["p"][0:3] = normalize_f32(d)
            if lt == "sun":
                rec["type"] = LIGHT_SUN
                angle = float(lj.get("angle", 0.533))                                    # degrees (light/sun.art:1)
                half = F(F(angle) / F(2)) / F(180) * F(3.14159265359)                    # rad(angle/2), core/common.art:20
                rec["p"][3] = F(math.cos(float(half)))
                sun_params[len(inf_l)] = (d, angle, "radiance" in lj, _color(lj["radiance"] if "radiance" in lj else lj.get("irradiance"), (1, 1, 1)))
                if "radiance" in lj:
                    rec["p"][4:7] = _color(lj["radiance"], (1, 1, 1))
                else:                                                                    # irradiance / sun_area_from_srad(rad(angle/2)), sun.art:5
                    rec["p"][4:7] = (_color(lj.get("irradiance"), (1, 1, 1)) * (F(1) / (F(3.14159265359) * half * half))).astype(F)
            else:
                rec["type"] = LIGHT_DIRECTIONAL
                rec["p"][3:6] = _color(lj.get("irradiance"), (1, 1, 1))
                sun_params[len(inf_l)] = (d, 0.0, False, rec["p"][3:6].copy())
            inf_l.append(rec)
        elif lt == "point":
            rec["type"] = LIGHT_POINT
            rec["p"][0:3] = _vec3(lj.get("position"), (0, 0, 0))
            if "power" in lj:
                rec["p"][3:6] = (_color(lj["power"], (0, 0, 0)) * F(1.0 / (4 * PI))).astype(F)
            else:
                rec["p"][3:6] = _color(lj.get("intensity"), (1, 1, 1))
            # PointLight.cpp:18-31: the cached colour is the power (intensity * 4 pi), flux = its mean
            flux = float(np.mean(_color(lj["power"], (0, 0, 0)) if "power" in lj else _color(lj.get("intensity"), (1, 1, 1)) * F(4 * PI)))
            fin_sel.append((rec["p"][0:3].copy(), None, flux))
            rec_raw = (_color(lj["power"], (0, 0, 0)), True) if "power" in lj else (_color(lj.get("intensity"), (1, 1, 1)), False)
            point_raw[id(rec)] = rec_raw
            fin_l.append(rec)
        elif lt == "spot":
            # SpotLight.cpp:11-20,62-90: cutoff / falloff in degrees; rad(x) = x / 180 * pi in f32 (core/common.art:20); the two
            # cosines are per-light constants, evaluated here (double, rounded once) exactly as the C++ recogniser does
            cut, fall = spot_cosines(float(lj.get("cutoff", 30.0)), float(lj.get("falloff", 20.0)))
            d = _vec3(lj.get("direction"), (0, 0, 1)).astype(np.float64)
            rec["type"] = LIGHT_SPOT
            rec["p"][0:3] = _vec3(lj.get("position"), (0, 0, 0))
            rec["p"][3:6] = (d / np.linalg.norm(d)).astype(F)
            rec["p"][6], rec["p"][7] = cut, fall
            if "power" in lj:   # spot_from_power, light/spot.art:1-6
                factor = F(F(2) * F(3.14159265359) * (F(1) - F(0.5) * F(fall) - F(0.5) * F(cut)))
                rec["p"][8:11] = (_color(lj["power"], (0, 0, 0)) * (F(1) / factor)).astype(F)
            else:
                rec["p"][8:11] = _color(lj.get("intensity"), (1, 1, 1))
            spot_angles[len(fin_l)] = (float(lj.get("cutoff", 30.0)), float(lj.get("falloff", 20.0)), "power" in lj, _color(lj["power"], (0, 0, 0)) if "power" in lj else None)
            # SpotLight.cpp:17-39: power_factor = 2 pi (1 - (cos cutoff + cos falloff) / 2)
            pf = 2 * PI * (1 - 0.5 * (math.cos(math.radians(float(lj.get("cutoff", 30.0)))) + math.cos(math.radians(float(lj.get("falloff", 20.0))))))
            flux = float(np.mean(_color(lj["power"], (0, 0, 0)) if "power" in lj else _color(lj.get("intensity"), (1, 1, 1)) * F(pf)))
            fin_sel.append((rec["p"][0:3].copy(), rec["p"][3:6].copy(), flux))
            fin_l.append(rec)
        elif lt == "area":
            ename = lj.get("entity", "")
            if ename not in ent_info:
                raise SceneError(f"No entity named '{ename}' exists for area light")
            ei = ent_info[ename]
            info = shape_info[ei["shape_id"]]
            t = ei["transform"]
            if info["type"] != SHAPE_TRIMESH:
                # analytic sphere: AreaLight.cpp:166-190 (RepresentationType::Sphere), light/area.art:260-316
                org, radius = info["origin"], float(info["radius"])
                area = ellipsoid_area(t[:3, :3].astype(F), radius)
                rec["type"] = LIGHT_SPHERE_AREA
                rec["entity_id"] = ei["id"]
                if "power" in lj:
                    rec["p"][0:3] = (_color(lj["power"], (0, 0, 0)) * F(1.0 / PI / area)).astype(F)
                else:
                    rec["p"][0:3] = _color(lj.get("radiance"), (1, 1, 1))
                rec["p"][3:6], rec["p"][6], rec["p"][7] = org, radius, area
                fin_of_entity[ename] = len(fin_l)
                col = _color(lj["power"], (0, 0, 0)) if "power" in lj else _color(lj.get("radiance"), (1, 1, 1)) * F(area * PI)   # AreaLight.cpp:100-113
                fin_sel.append(((t[:3, :3] @ np.asarray(org, np.float64) + t[:3, 3]).astype(F), None, float(np.mean(col))))
                fin_l.append(rec)
                continue
            plane = info["plane"] if lj.get("optimize", True) else None
            if plane is not None:
                origin = (t[:3, :3] @ plane["origin"].astype(np.float64) + t[:3, 3]).astype(F)
                xa = (t[:3, :3] @ plane["x_axis"].astype(np.float64)).astype(F)
                ya = (t[:3, :3] @ plane["y_axis"].astype(np.float64)).astype(F)
                cr = np.cross(xa.astype(np.float64), ya.astype(np.float64))
                area = float(np.linalg.norm(cr))
                rec["type"] = LIGHT_PLANE_AREA
                rec["p"][0:3], rec["p"][3:6], rec["p"][6:9] = origin, xa, ya
                rec["p"][9:12] = (cr / area).astype(F)
                rec["p"][12] = area
                rec["p"][13:21] = plane["texcoords"].reshape(-1)
                rad_off = 21
                sel_pos, sel_dir = (origin + xa * F(0.5) + ya * F(0.5)).astype(F), (cr / area).astype(F)   # AreaLight.cpp:69-70
            else:
                mesh = info["mesh"]
                d = (info["hi"] - info["lo"]).astype(np.float64)
                w = np.linalg.norm(t[:3, 0] * d[0])
                h = np.linalg.norm(t[:3, 1] * d[1])
                dd = np.linalg.norm(t[:3, 2] * d[2])
                half = d[0] * d[1] + d[0] * d[2] + d[1] * d[2]
                area = mesh.compute_area() * (w * h + w * dd + h * dd) / half
                rec["type"] = LIGHT_SHAPE_AREA
                rad_off = 0
                ctr = (info["lo"].astype(np.float64) + info["hi"].astype(np.float64)) / 2
                sel_pos, sel_dir = (t[:3, :3] @ ctr + t[:3, 3]).astype(F), None                              # AreaLight.cpp:89-90
            rec["entity_id"] = ei["id"]
            if "power" in lj:
                rec["p"][rad_off:rad_off + 3] = (_color(lj["power"], (0, 0, 0)) * F(1.0 / PI / area)).astype(F)
            else:
                rec["p"][rad_off:rad_off + 3] = _color(lj.get("radiance"), (1, 1, 1))
            fin_of_entity[ename] = len(fin_l)
            col = _color(lj["power"], (0, 0, 0)) if "power" in lj else _color(lj.get("radiance"), (1, 1, 1)) * F(area * PI)
            fin_sel.append((sel_pos, sel_dir, float(np.mean(col))))
            fin_l.append(rec)
        else:
            raise SceneError(f"light type '{lt}' is outside the supported path (SURVEY §8f)")

    # ---- embedded light tables (LoaderLight.h:27, LoaderLight.cpp:316-420): once >= 10 finite lights are "simple" (constant parameters),
    # they are written into per-class fix-tables and come FIRST in the id space, class by class (the reference iterates an
    # unordered_map of class names: implementation-defined; here classes in order of first appearance), the other lights after them.
    # A simple point light's table entry holds power / (4 pi), the power being intensity * 4 pi (PointLight.cpp:18-31,63-70).
    embed_class = []
    for rec in fin_l:
        t_ = int(rec["type"])
        embed_class.append({LIGHT_POINT: "SimplePointLight", LIGHT_SPOT: "SimpleSpotLight", LIGHT_PLANE_AREA: "SimplePlaneLight",
                            LIGHT_SHAPE_AREA: "SimpleAreaLight", LIGHT_SPHERE_AREA: "SimpleSphereLight"}.get(t_))
    embedded = sum(c is not None for c in embed_class) >= 10
    if embedded:
        classes = []
        for c in embed_class:
            if c is not None and c not in classes:
                classes.append(c)
        order = [i for c in classes for i, ec in enumerate(embed_class) if ec == c] + [i for i, ec in enumerate(embed_class) if ec is None]
        remap = {old: new for new, old in enumerate(order)}
        fin_l = [fin_l[i] for i in order]
        fin_sel = [fin_sel[i] for i in order]
        spot_angles = {remap[k]: v for k, v in spot_angles.items()}
        fin_of_entity = {k: remap[v] for k, v in fin_of_entity.items()}
        sr = F(4 * PI)
        for rec in fin_l:
            if int(rec["type"]) == LIGHT_POINT:
                colour, is_power = point_raw[id(rec)]
                rec["p"][3:6] = ((colour if is_power else (colour * sr).astype(F)) / sr).astype(F)

    # ---- light selector (LoaderLight.cpp:423-452): one light or none -> uniform whatever was asked for
    selector_data = np.zeros(0, F)
    if sel in ("uniform", "") or len(inf_l) + len(fin_l) <= 1 or not fin_l:
        technique["light_selector"] = SELECTOR_UNIFORM
    elif sel == "simple":
        technique["light_selector"] = SELECTOR_CDF
        selector_data = light_cdf([f for _, _, f in fin_sel])
    elif sel == "hierarchy":
        technique["light_selector"] = SELECTOR_HIERARCHY
        selector_data = light_hierarchy(fin_sel)
    else:
        technique["light_selector"] = SELECTOR_UNIFORM   # LoaderLight.cpp:448-450: anything else is the uniform selector

    # ---- materials
    materials = np.zeros(len(groups), MATERIAL_DTYPE)
    for mid, (bname, emis) in enumerate(mat_keys):
        bj = bsdfs[bname]
        bt = bj.get("type", "").lower()
        rec = materials[mid]
        rec["light_id"] = fin_of_entity[emis] if emis is not None else -1
        rec["tex"] = (-1, -1)
        rec["map_tex"] = -1
        if bt in ("bumpmap", "normalmap"):
            # MapBSDF.cpp:17-55: a wrapper that replaces the shading frame of its inner BSDF (bsdf/map.art:39-68)
            inner = bj.get("bsdf", "")
            if inner not in bsdfs or bsdfs[inner].get("type", "").lower() in ("bumpmap", "normalmap"):
                raise SceneError(f"bsdf '{bname}': missing inner bsdf / nested maps are outside the supported path")
            mp = bj.get("map")
            if not (isinstance(mp, str) and mp.strip() in tex_json):
                raise SceneError(f"bsdf '{bname}': the map must be a texture name")
            rec["map_kind"] = MAP_BUMP if bt == "bumpmap" else MAP_NORMAL
            rec["map_tex"] = texture_id(mp.strip())
            rec["map_strength"] = float(bj.get("strength", 1.0))
            bj = bsdfs[inner]
            bt = bj.get("type", "").lower()
        if bt in ("diffuse", "roughdiffuse"):
            alpha = bj.get("alpha", bj.get("roughness", 0.0))
            if isinstance(alpha, str) or float(alpha) > 1.1920928955e-07:
                raise SceneError("Oren-Nayar (rough) diffuse is outside the supported path")
            rec["bsdf"] = BSDF_DIFFUSE
            rec["p"][0:3], rec["tex"][0] = color_or_tex(bj.get("reflectance"), (0.8, 0.8, 0.8))
        elif bt in ("dielectric", "glass", "roughdielectric", "thindielectric"):
            rough = [bj.get(k, 0) or 0 for k in ("roughness", "alpha", "roughness_u", "roughness_v", "alpha_u", "alpha_v")]
            if bj.get("thin", False) or any(isinstance(r, str) or float(r) > 0 for r in rough):
                raise SceneError("thin / rough dielectrics are outside the supported path")
            iors = {"vacuum": 1.0, "bk7": 1.5046}
            ext = bj.get("ext_ior", iors.get(str(bj.get("ext_ior_material", "")).lower(), 1.0))
            int_ = bj.get("int_ior", iors.get(str(bj.get("int_ior_material", "")).lower(), 1.5046))
            rec["bsdf"] = BSDF_DIELECTRIC
            rec["p"][0], rec["p"][1] = float(ext), float(int_)
            rec["p"][2:5], rec["tex"][0] = color_or_tex(bj.get("specular_reflectance"), (1, 1, 1))
            rec["p"][5:8], rec["tex"][1] = color_or_tex(bj.get("specular_transmittance"), (1, 1, 1))
        elif bt in ("mirror", "conductor", "roughconductor"):
            # ConductorBSDF.cpp:13-35 + BSDF::setupRoughness (BSDF.cpp:53-98): no roughness property -> make_delta_distribution, else
            # make_vndf_ggx_distribution over compute_explicit(roughness, anisotropic) or the explicit (roughness_u, roughness_v)
            old = any(k in bj for k in ("alpha", "alpha_u", "alpha_v"))
            rp = "alpha" if old else "roughness"
            if any(k in bj for k in (rp, rp + "_u", rp + "_v")):
                if str(bj.get("distribution", "vndf_ggx")).lower() in ("ggx", "beckmann"):
                    raise SceneError("the plain ggx / beckmann distributions are outside the supported path (vndf_ggx is the default)")
                vals = [bj.get(k) for k in (rp, rp + "_u", rp + "_v", "anisotropic")]
                if any(isinstance(x, str) for x in vals):
                    raise SceneError("textured roughness is outside the supported path")
                if (rp + "_u") in bj or (rp + "_v") in bj:
                    au, av = F(bj.get(rp + "_u", 0.1)), F(bj.get(rp + "_v", 0.1))
                else:   # microfacet::compute_explicit, core/microfacet.art:427-432
                    r, an = F(bj.get(rp, 0.1)), F(bj.get("anisotropic", 0.0))
                    aspect = F(1) if an == 0 else F(np.sqrt(F(1) - F(min(max(float(an), 0.0), 1.0)) * F(0.99)))
                    au, av = F(r / aspect), F(r * aspect)
                rec["distribution"], rec["alpha_u"], rec["alpha_v"] = MICROFACET_VNDF_GGX, au, av
            conductors = {"none": ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), "aluminum": ((1.34560, 0.96521, 0.61722), (7.47460, 6.39950, 5.30310)),
                          "brass": ((0.44400, 0.52700, 1.09400), (3.69500, 2.76500, 1.82900)), "copper": ((0.27105, 0.67693, 1.31640), (3.60920, 2.62480, 2.29210)),
                          "gold": ((0.18299, 0.42108, 1.37340), (3.4242, 2.34590, 1.77040)), "iron": ((2.91140, 2.94970, 2.58450), (3.08930, 2.93180, 2.76700)),
                          "lead": ((1.91000, 1.83000, 1.44000), (3.51000, 3.40000, 3.18000)), "mercury": ((2.07330, 1.55230, 1.06060), (5.33830, 4.65100, 3.86280)),
                          "platinum": ((2.37570, 2.08470, 1.84530), (4.26550, 3.71530, 3.13650)), "silver": ((0.15943, 0.14512, 0.13547), (3.92910, 3.19000, 2.38080)),
                          "titanium": ((2.74070, 2.54180, 2.26700), (3.81430, 3.43450, 3.03850))}   # BSDF.cpp:29-42
            d_eta, d_k = conductors.get(str(bj.get("material", "")).lower(), conductors["none"])
            eta, kk = _color(bj.get("eta"), d_eta), _color(bj.get("k"), d_k)
            rec["bsdf"] = BSDF_CONDUCTOR
            rec["p"][0:3], rec["p"][3:6] = eta, kk
            rec["p"][6:9], rec["tex"][0] = color_or_tex(bj.get("specular_reflectance"), (1, 1, 1))
            # conductor.art:133-135: the mirror is chosen when eta and k are constants the generator printed into the text, i.e.
            # (default specialisation, ShadingTree.cpp:868-884) eta black and k white
            rec["p"][9] = 1.0 if (np.all(np.abs(eta) <= 1.1920929e-07) and np.all(np.abs(kk - 1) <= 1.1920929e-07)) else 0.0
        else:
            raise SceneError(f"bsdf type '{bt}' is outside the supported path (SURVEY §8f)")

    # ---- camera (PerspectiveCamera.cpp:11-107, Camera.cpp:5-15)
    cj = doc.get("camera", {"type": "perspective"})
    if cj.get("type", "perspective") != "perspective":
        raise SceneError(f"camera type '{cj.get('type')}' is outside the supported path")
    if float(cj.get("aperture_radius", 0)) > 1.1920928955e-07:
        raise SceneError("depth of field is outside the supported path")
    cam = np.zeros((), CAMERA_DTYPE)
    if "vfov" in cj:
        vertical, fov = 1, float(cj["vfov"])
    elif "hfov" in cj:
        vertical, fov = 0, float(cj["hfov"])
    else:
        vertical, fov = 0, float(cj.get("fov", 60.0))
    # the generator streams the value into shader text with 6 significant digits
    cam["fov"] = float("%g" % float(F(fov) * F(DEG2RAD)))
    cam["fov_vertical"] = vertical
    cam["aspect"] = float("%f" % float(cj["aspect_ratio"])) if "aspect_ratio" in cj else 0.0
    near = float(cj.get("near_clip", 0.0))
    far = float(cj.get("far_clip", np.finfo(np.float32).max))
    if far < near:
        near, far = far, near
    cam["tmin"], cam["tmax"] = float("%g" % near), float("%g" % far)
    if "transform" in cj:
        t = parse_transform(cj["transform"])
        cam["eye"] = t[:3, 3].astype(F)
        cam["dir"] = t[:3, 2].astype(F)
        cam["up"] = t[:3, 1].astype(F)
    elif n_ent == 0:
        cam["eye"], cam["dir"], cam["up"] = (0, 0, 0), (0, 0, -1), (0, 1, 0)
    else:
        aspect = float(cam["aspect"]) if cam["aspect"] > 0 else fw / fh
        d3 = (bb_hi - bb_lo).astype(np.float64)
        a = d3[0] / (2 * (aspect if vertical else 1))
        b = d3[1] / (2 * (aspect if not vertical else 1))
        s = math.sin(float(cam["fov"]) / 2)
        d = 0.0 if abs(s) <= 1.1920928955e-07 else max(a, b) * math.sqrt(1 / (s * s) - 1)
        c = (bb_lo.astype(np.float64) + bb_hi.astype(np.float64)) / 2
        cam["eye"], cam["dir"], cam["up"] = (c[0], c[1], bb_hi[2] + d), (0, 0, -1), (0, 1, 0)

    if n_ent == 0:
        bb_lo = np.zeros(3, F)
        bb_hi = np.zeros(3, F)

    def _arr(lst):
        return np.asarray(lst, LIGHT_DTYPE) if lst else np.zeros(0, LIGHT_DTYPE)

    lk = np.zeros(len(lookups), LOOKUP_DTYPE)
    for i, (ty, fl, off) in enumerate(lookups):
        lk[i] = (ty, fl, off)
    return SceneTables(
        entities=np.ascontiguousarray(ent_table), shape_lookups=lk,
        shape_data=np.frombuffer(bytes(blob), np.uint8).copy(), leaves=leaves,
        entity_per_material=np.asarray([len(g) for g in groups], np.int32), materials=materials,
        infinite_lights=_arr(inf_l), finite_lights=_arr(fin_l), camera=cam, technique=technique,
        bbox_min=bb_lo.astype(F), bbox_max=bb_hi.astype(F), film_size=(fw, fh),
        embedded_lights=embedded, selector_data=selector_data, textures=(np.asarray(tex_recs, TEXTURE_DTYPE) if tex_recs else np.zeros(0, TEXTURE_DTYPE)), images=images,
        image_files=image_files,
        aux_data=(np.concatenate(aux_words).astype(F) if aux_words else np.zeros(0, F)), texture_names=list(tex_ids), entity_names=names, material_names=[f"{b}{'@' + e if e else ''}" for b, e in mat_keys], spot_angles=spot_angles, sun_params=sun_params)
