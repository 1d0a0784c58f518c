"""Stage scripts + registries as the reference's runtime would hand them to a device plugin.

The reference's `Runtime` never gives a device descriptors: it gives it *program text* per pipeline stage plus parameter
registries (`TechniqueVariantShaderSet` + `LocalRegistry`, src/runtime/technique/TechniqueVariant.h:12-35;
`ICompilerDevice::compileAndGet(script, function)`, src/runtime/Runtime.cpp:631-657). The B200 plugin's compiler device
(`csrc/host/script_recognizer.cpp`) parses that text back into descriptors. To test it without the reference's loader
(which cannot be built here) this module RECONSTRUCTS the text from `SceneTables`, statement by statement
after the generators:

  stage prologue / database / scene     src/runtime/shader/ShaderUtils.cpp:13-53,55-62,101-141,174-204,206-216
  ig_ray_generation_shader              src/runtime/shader/RayGenerationShader.cpp:12-70, camera/PerspectiveCamera.cpp:26-67
  ig_hit_shader / ig_miss_shader        src/runtime/shader/HitShader.cpp:16-53, MissShader.cpp:15-47
  BSDFs                                 src/runtime/bsdf/DiffuseBSDF.cpp:13-27, DielectricBSDF.cpp:13-41, ConductorBSDF.cpp:13-35, BSDF.cpp:53-98
                                        (setupRoughness), MapBSDF.cpp:17-55 (bumpmap / normalmap)
  textures                              src/runtime/pattern/CheckerBoardPattern.cpp:13-33, ImagePattern.cpp:15-73 (8-bit files: resource id in the
                                        LocalRegistry + device.load_packed_image_by_id), loader/ShadingTree.cpp:375-405,795-840 and
                                        Transpiler.cpp:991,1301 (a texture name inside a colour -> `vec4_to_color(color_to_vec4(tex_<id>(ctx)))`);
                                        float images travel as OpenEXR files (device.load_image_by_id)
  textured environments, sky            src/runtime/light/EnvironmentLight.cpp:40-110, SkyLight.cpp:49-76, LoaderUtils.cpp:108-132 (cdf file),
                                        CDF.cpp:71-151
  lights + tables                       src/runtime/light/{AreaLight.cpp:115-220, PointLight.cpp:44-62, EnvironmentLight.cpp:103-110},
                                        src/runtime/loader/LoaderLight.cpp:106-247,423-453; embedded simple point lights
                                        (>= 10 simple lights, LoaderLight.h:27): LoaderLight.cpp:171-236,397-422, PointLight.cpp:64-78
  technique                             src/runtime/technique/PathTechnique.cpp:35-79
  value inlining rules                  src/runtime/loader/ShadingTree.cpp:868-981 (default specialisation: only all-zero / all-one
                                        values are printed into the text, with std::to_string's 6 decimals; everything else goes
                                        to the stage's LocalRegistry at full precision)

It is a reconstruction, not a captured `--dump-shader` output: whitespace may differ, the grammar does not.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .scene import (BSDF_CONDUCTOR, BSDF_DIELECTRIC, BSDF_DIFFUSE, LIGHT_ENV_CONST, LIGHT_PLANE_AREA, LIGHT_POINT, LIGHT_SHAPE_AREA, LIGHT_SPHERE_AREA, LIGHT_SPOT,
                    IMAGE_RGBA32F, LIGHT_ENV_TEX, LIGHT_ENV_TEXTURED, MAP_BUMP, MAP_NONE, MICROFACET_VNDF_GGX, SHAPE_SPHERE, TEX_CHECKERBOARD, SceneTables)

STD_LIB_STUB = "// <the Artic standard library (ig_api[], ScriptCompiler.cpp:36-51) precedes every stage>\nfn @make_dummy() = 0;\n\n"


@dataclass
class Registry:
    """ParameterSet, src/runtime/ParameterSet.h:6-12"""
    ints: dict = field(default_factory=dict)
    floats: dict = field(default_factory=dict)
    vectors: dict = field(default_factory=dict)
    colors: dict = field(default_factory=dict)


@dataclass
class Stage:
    function: str
    script: str
    local: Registry


@dataclass
class StageSet:
    raygen: Stage
    miss: Stage
    hits: list
    global_registry: Registry
    fix_tables: dict = field(default_factory=dict)   # what LoaderLight::embedLights adds to the scene database: class name -> (n, words) float32
    resource_map: list = field(default_factory=list)  # IRenderDevice::SceneSettings::resource_map: file of resource id k


def _ts(x: float) -> str:
    return f"{float(x):.6f}"   # std::to_string(float)


def _stream(x: float) -> str:
    return f"{float(x):.6g}"   # operator<<(ostream&, float), default precision


def write_exr(path: str, rgb: np.ndarray) -> None:
    """(H, W, 3) float32, rows top-down -> an uncompressed single-part scanline OpenEXR file with FLOAT channels B, G, R: the simplest file the
    format allows (what the reference's loader would find on disk is whatever wrote the user's environment map; the host layer's reader is
    checked against every compression separately, tests/test_plugin_host.py::test_host_exr_reader)."""
    import struct
    rgb = np.ascontiguousarray(rgb, np.float32)
    h, w, _ = rgb.shape

    def attr(name, ty, body):
        return name.encode() + b"\0" + ty.encode() + b"\0" + struct.pack("<i", len(body)) + body
    chlist = b"".join(c + b"\0" + struct.pack("<iB3xii", 2, 0, 1, 1) for c in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<4i", 0, 0, w - 1, h - 1)
    head = (struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0") + attr("dataWindow", "box2i", box)
            + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
            + attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    line = 8 + 12 * w
    first = len(head) + 8 * h
    with open(path, "wb") as fh:
        fh.write(head + b"".join(struct.pack("<Q", first + y * line) for y in range(h)))
        for y in range(h):
            fh.write(struct.pack("<ii", y, 12 * w) + rgb[y, :, 2].tobytes() + rgb[y, :, 1].tobytes() + rgb[y, :, 0].tobytes())


class _Tree:
    """ShadingTree: closure ids + the inline-or-registry decision."""

    def __init__(self, local: Registry, specialization: str):
        self.local, self.mode, self.ids, self.header, self.textures = local, specialization, {}, [], set()
        self.resources: list = []   # LoaderContext::registerExternalResource: shared by all stages of a scene (set by generate())
        self.cache_dir = None       # the loader's cache directory: float images (EXR) and cdf buffers are files there

    def resource(self, path: str) -> int:
        if path not in self.resources:
            self.resources.append(path)
        return self.resources.index(path)

    def closure(self, name: str) -> str:
        return str(self.ids.setdefault(name, len(self.ids)))

    def _embed(self, vals, zero=True, one=True) -> bool:
        if self.mode == "force":
            return True
        if self.mode == "disable":
            return False
        v = np.asarray(vals, np.float32).ravel()
        return bool((zero and np.all(np.abs(v) <= 1.1920929e-07)) or (one and np.all(np.abs(v - 1) <= 1.1920929e-07)))

    def number(self, cid, prop, val, dynamic=False, zero=True, one=True) -> str:
        if not dynamic and self._embed([val], zero, one):
            return _ts(val)
        key = f"{cid}_{prop}"
        self.local.floats[key] = float(np.float32(val))
        self.header.append(f'  let var_num_{key} = registry::get_local_parameter_f32("{key}", 0);\n')
        return f"var_num_{key}"

    def integer(self, cid, prop, val) -> str:
        key = f"{cid}_{prop}"
        self.local.ints[key] = int(val)
        self.header.append(f'  let var_int_{key} = registry::get_local_parameter_i32("{key}", 0);\n')
        return f"var_int_{key}"

    def color(self, cid, prop, rgb, zero=True, one=True) -> str:
        if self._embed(rgb, zero, one):
            return f"make_color({_ts(rgb[0])}, {_ts(rgb[1])}, {_ts(rgb[2])}, 1)"
        key = f"{cid}_{prop}"
        self.local.colors[key] = tuple(float(np.float32(c)) for c in rgb) + (1.0,)
        self.header.append(f'  let var_color_{key} = registry::get_local_parameter_color("{key}", color_builtins::black);\n')
        return f"var_color_{key}"

    def vector(self, cid, prop, xyz, dynamic=False) -> str:
        if not dynamic and self._embed(xyz):
            return f"make_vec3({_ts(xyz[0])}, {_ts(xyz[1])}, {_ts(xyz[2])})"
        key = f"{cid}_{prop}"
        self.local.vectors[key] = tuple(float(np.float32(c)) for c in xyz)
        self.header.append(f'  let var_vec_{key} = registry::get_local_parameter_vec3("{key}", vec3_expand(0));\n')
        return f"var_vec_{key}"

    def texture(self, t: SceneTables, tex: int) -> str:
        """The look-up a texture name turns into inside a colour expression; the texture's own `let` goes to the header once per stage
        (ShadingTree::registerTextureUsage, ShadingTree.cpp:795-804)."""
        cid = self.closure(f"__tex_{tex}")
        if tex not in self.textures:
            self.textures.add(tex)
            rec = t.textures[tex]
            a, b, c, d, e, f = (float(x) for x in rec["transform"])
            if (a, b, c, d, e, f) == (1, 0, 0, 0, 1, 0):
                tr = "mat3x3_identity()"
            else:   # LoaderUtils::inlineMatrix: columns, streamed with the default precision
                tr = "make_mat3x3(" + ", ".join("make_vec3(%s, %s, %s)" % tuple(_stream(x) for x in col) for col in ((a, d, 0), (b, e, 0), (c, f, 1))) + ")"
            if int(rec["type"]) == TEX_CHECKERBOARD:
                p = rec["p"]
                c0, c1 = self.color(cid, "color0", p[2:5]), self.color(cid, "color1", p[5:8])
                sx, sy = self.number(cid, "scale_x", p[0]), self.number(cid, "scale_y", p[1])
                self.header.append(f"  let tex_{cid} : Texture = make_checkerboard_texture(make_vec2({sx}, {sy}), {c0}, {c1}, {tr});\n")
            else:   # ImagePattern.cpp:15-73: the file travels as a resource id in the stage's LocalRegistry, the device loads it
                fmt, arr = t.images[int(rec["image"])]
                ref = t.image_files[int(rec["image"])]
                if ref is None:
                    raise ValueError("the sky texture is written by its light (SkyLight.cpp:28-47), not referenced as a texture")
                file = ref[0] if fmt != IMAGE_RGBA32F else self.float_image_file(t, int(rec["image"]))
                self.local.ints[f"img_{cid}"] = self.resource(file)
                wraps = ["make_repeat_border()", "make_clamp_border()", "make_mirror_border()"]
                wu, wv = wraps[int(rec["border_u"])], wraps[int(rec["border_v"])]
                wrap = wu if wu == wv else f"make_split_border({wu}, {wv})"
                filt = ["make_nearest_filter()", "make_bilinear_filter()", "make_bicubic_filter()"][int(rec["filter"])]
                channels = 1 if arr.ndim == 2 else 4   # Image::loadResolution(filename).Channels == 1 ? 1 : 4
                load = (f"device.load_image_by_id(img_{cid}_res_id, {channels})" if fmt == IMAGE_RGBA32F      # Image::isPacked: 8-bit formats only
                        else f"device.load_packed_image_by_id(img_{cid}_res_id, {channels}, {'true' if ref[1] else 'false'})")
                self.header.append(f'  let img_{cid}_res_id = registry::get_local_parameter_i32("img_{cid}", 0);\n'
                                   f"  let img_{cid} = {load};\n"
                                   f"  let tex_{cid} : Texture = make_image_texture({wrap}, {filt}, img_{cid}, {tr});\n")
        return f"vec4_to_color(color_to_vec4(tex_{cid}(ctx)))"

    def float_image_file(self, t: SceneTables, image: int, stem: str | None = None) -> str:
        """The OpenEXR file a float image of the tables stands for: the fixtures are .npy / .npz arrays (or the baked sky), the reference's loader
        and device deal in files, so the pixels are written into the cache directory as the simplest EXR there is."""
        import os
        if self.cache_dir is None:
            raise ValueError("float images need a cache directory to live in as files (refscript.generate(cache_dir=...))")
        path = os.path.join(self.cache_dir, (stem or f"image_{image}") + ".exr")
        if not os.path.exists(path):
            write_exr(path, t.images[image][1][::-1, :, :3])   # the tables hold rows bottom-up (Image::flipY); the file is top-down
        return path

    def cdf_file(self, t: SceneTables, light) -> tuple:
        """(file, size_x, size_y) of a textured environment light's 2-D cdf as CDF::computeForImage writes it (CDF.cpp:71-151): marginal, then the
        conditional rows, raw floats."""
        import os
        if self.cache_dir is None:
            raise ValueError("environment cdfs need a cache directory (refscript.generate(cache_dir=...))")
        first, sx, sy = (int(x) for x in light["p"][13:16].view(np.int32))
        path = os.path.join(self.cache_dir, f"cdf_{first}.bin")
        if not os.path.exists(path):
            np.ascontiguousarray(t.aux_data[first:first + sy + sy * sx], np.float32).tofile(path)
        return path, sx, sy

    def pull_header(self) -> str:
        h, self.header = "".join(self.header), []
        return h


def _prologue(t: SceneTables) -> str:
    has_sphere = any(int(lk["type_id"]) == SHAPE_SPHERE for lk in t.shape_lookups)
    s = ("  let spi = settings.spi;\n"
         "  let render_config = make_render_config_from_settings(settings, spi);\n"
         "  let device = make_nvvm_device(settings.device, render_config, make_default_gpu_kernel_config());\n"
         "  let payload_info = PayloadInfo{ primary_count = 6, secondary_count = 0 };\n"
         '  let scene_bbox = make_bbox(registry::get_global_parameter_vec3("__scene_bbox_lower", vec3_expand(0)), '
         'registry::get_global_parameter_vec3("__scene_bbox_upper", vec3_expand(0))); maybe_unused(scene_bbox);\n\n')
    return s, has_sphere


def _database(has_sphere: bool) -> str:
    s = ("  let entities = load_entity_table(device); maybe_unused(entities);\n"
         "  let shapes_trimesh = load_shape_table(device, @|_, data| { \n    make_trimesh_shape(load_trimesh(data))\n  });\n  maybe_unused(shapes_trimesh);\n")
    if has_sphere:
        s += ("  let shapes_sphere = load_shape_table(device, @|_, data| { \n    make_sphere_shape(load_sphere(data))\n  });\n  maybe_unused(shapes_sphere);\n"
              "  let shapes = load_shape_table(device, @|type_id, data| { match type_id {\n    0 => make_trimesh_shape(load_trimesh(data)),\n"
              "    _ => make_sphere_shape(load_sphere(data))\n  }});\n")
    else:
        s += "  let shapes = shapes_trimesh;\n"
    s += "  maybe_unused(shapes);\n"
    s += ('  let scene  = Scene {\n    num_entities  = registry::get_global_parameter_i32("__entity_count", 0),\n'
          '    num_materials = registry::get_global_parameter_i32("__material_count", 0),\n    shapes   = shapes,\n    entities = entities,\n  };\n')
    return s


def _lights(t: SceneTables, tree: _Tree) -> str:
    s = ""
    bbox = ('make_bbox(registry::get_global_parameter_vec3("__scene_bbox_lower", vec3_expand(0)), '
            'registry::get_global_parameter_vec3("__scene_bbox_upper", vec3_expand(0)))')
    inf_names, fin_names = [], []
    for i, l in enumerate(t.infinite_lights):
        cid = tree.closure(f"__inf_light_{i}")
        if int(l["type"]) in (6, 7):   # SunLight.cpp:28-57, DirectionalLight.cpp:25-41 (the scene box is the shader's `scene_bbox`, LoaderUtils.cpp:10-19)
            raw_dir, angle, use_radiance, colour = t.sun_params[i]
            if int(l["type"]) == 6:
                col = tree.color(cid, "radiance" if use_radiance else "irradiance", colour)
                ang = tree.number(cid, "angle", angle)
                direction = tree.vector(cid, "direction", raw_dir)
                rad = col if use_radiance else f"color_mulf({col}, 1 / sun_area_from_srad(rad({ang}/2)))"
                s += tree.pull_header() + f"  let light_{cid} = make_sun_light({i}, vec3_normalize({direction}), scene_bbox, math_builtins::cos(rad({ang}/2)), {rad}, false);\n"
            else:
                col = tree.color(cid, "irradiance", colour)
                direction = tree.vector(cid, "direction", raw_dir)
                s += tree.pull_header() + f"  let light_{cid} = make_directional_light({i}, vec3_normalize({direction}), scene_bbox, {col});\n"
            inf_names.append(f"light_{cid}")
            continue
        if int(l["type"]) in (LIGHT_ENV_TEXTURED, LIGHT_ENV_TEX):
            tex = int(l["p"][12:13].view(np.int32)[0])
            scale = tree.color(cid, "scale", l["p"][0:3])
            tr = "make_mat3x3(" + ",".join(tree.vector(cid, f"_transform_c{k}", l["p"][3 + 3 * k:6 + 3 * k]) for k in range(3)) + ")"   # ShadingTree::getInlineMatrix3
            image = int(t.textures[tex]["image"])
            if t.image_files[image] is None:   # SkyLight.cpp:49-76: the baked sky model, its texture and cdf bound right here, resources by literal id
                file = tree.float_image_file(t, image, f"skytex_{i}")
                cdf, sx, sy = tree.cdf_file(t, l)
                s += tree.pull_header() + (f"  let sky_tex_{cid} = make_image_texture(make_repeat_border(), make_bilinear_filter(), device.load_image_by_id({tree.resource(file)}, 4), mat3x3_identity());\n"
                                           f"  let sky_cdf_{cid} = cdf::make_cdf_2d_from_buffer(device.load_buffer_by_id({tree.resource(cdf)}), {sx}, {sy});\n"
                                           f"  let light_{cid}   = make_environment_light_textured({i}, {bbox}, {scale}, sky_tex_{cid}, sky_cdf_{cid}, {tr});\n")
            else:                              # EnvironmentLight.cpp:40-110: radiance is a texture (ShadingTree::addTexture, the string case)
                rad = f"@|ctx:ShadingContext|->Color{{maybe_unused(ctx); {tree.texture(t, tex)}}}"
                if int(l["type"]) == LIGHT_ENV_TEXTURED:
                    cdf, sx, sy = tree.cdf_file(t, l)
                    s += tree.pull_header() + (f"  let cdf_{cid}   = cdf::make_cdf_2d_from_buffer(device.load_buffer_by_id({tree.resource(cdf)}), {sx}, {sy});\n"
                                               f"  let light_{cid} = make_environment_light_textured({i}, {bbox}, {scale}, {rad}, cdf_{cid}, {tr});\n")
                else:
                    s += tree.pull_header() + f"  let light_{cid} = make_environment_light({i}, {bbox}, {scale}, {rad}, {tr});\n"
            inf_names.append(f"light_{cid}")
            continue
        assert int(l["type"]) == LIGHT_ENV_CONST
        scale = tree.color(cid, "scale", (1, 1, 1))
        rad = "make_constant_texture(" + tree.color(cid, "radiance", l["p"][0:3]) + ")"
        ident = "make_mat3x3(make_vec3(1.000000, 0.000000, 0.000000),make_vec3(0.000000, 1.000000, 0.000000),make_vec3(0.000000, 0.000000, 1.000000))"
        s += tree.pull_header() + f"  let light_{cid} = make_environment_light({i}, {bbox}, {scale}, {rad}, {ident});\n"
        inf_names.append(f"light_{cid}")
    s += f"  let infinite_lights = LightTable {{\n    count = {len(inf_names)},\n    get   = @|id:i32| {{\n      match(id) {{\n"
    for i, n in enumerate(inf_names):
        s += f"      {i} => {n},\n"
    s += "      _ => make_null_light(id)\n    }\n  }};\n  maybe_unused(infinite_lights);\n"
    # LoaderLight.cpp:152-236: with >= 10 simple lights (LoaderLight.h:27) those come from per-class fix-tables and lead the id space, class by
    # class (ignis_b200/scene.py orders the records that way); only the others are written into the text
    embedded = []   # (class, count, offset)
    if t.embedded_lights:
        cls_of = {LIGHT_POINT: "SimplePointLight", LIGHT_SPOT: "SimpleSpotLight", LIGHT_PLANE_AREA: "SimplePlaneLight", LIGHT_SHAPE_AREA: "SimpleAreaLight",
                  LIGHT_SPHERE_AREA: "SimpleSphereLight"}
        for i, l in enumerate(t.finite_lights):
            c = cls_of.get(int(l["type"]))
            if c is None:
                break
            if c != "SimplePointLight":
                raise ValueError(f"embedded lights of class {c} are outside the script path (SimplePointLight only)")
            if embedded and embedded[-1][0] == c:
                embedded[-1] = (c, embedded[-1][1] + 1, embedded[-1][2])
            else:
                embedded.append((c, 1, i))
    n_embedded = sum(e[1] for e in embedded)
    for i, l in enumerate(t.finite_lights):
        if i < n_embedded:
            continue
        cid = tree.closure(f"__fin_light_{i}")
        ty, p = int(l["type"]), l["p"]
        if ty == LIGHT_POINT:
            origin = tree.vector(cid, "origin", p[0:3])
            inten = tree.color(cid, "intensity", p[3:6])
            s += tree.pull_header() + f"  let light_{cid} = make_point_light({i}, {origin}, {inten});\n"
        elif ty == LIGHT_PLANE_AREA:
            pre = f"ae_{cid}"
            rad = tree.color(cid, "radiance", p[21:24])
            args = [tree.vector(cid, pre + "_origin", p[0:3], dynamic=True), tree.vector(cid, pre + "_tangent", p[3:6]),
                    tree.vector(cid, pre + "_bitangent", p[6:9]), tree.vector(cid, pre + "_normal", p[9:12]), tree.number(cid, pre + "_area", p[12])]
            tcs = [tree.vector(cid, f"{pre}_t{k}", (p[13 + 2 * k], p[14 + 2 * k], 0), dynamic=True) for k in range(4)]
            s += tree.pull_header() + f"  let ae_{cid} = make_plane_area_emitter({', '.join(args)}, " + ", ".join(f"vec3_to_2({x})" for x in tcs) + ");\n"
            s += f"  let light_{cid} = make_area_light({i}, ae_{cid}, @|ctx| {{ maybe_unused(ctx); {rad} }});\n"
        elif ty == LIGHT_SHAPE_AREA:
            rad = tree.color(cid, "radiance", p[0:3])
            ent = tree.integer(cid, f"ae_{cid}_ent_id", int(l["entity_id"]))
            s += tree.pull_header() + f"  let ae_{cid} = make_shape_area_emitter_proxy({ent}, entities, shapes_trimesh);\n"
            s += f"  let light_{cid} = make_area_light({i}, ae_{cid}, @|ctx| {{ maybe_unused(ctx); {rad} }});\n"
        elif ty == LIGHT_SPOT:   # SpotLight.cpp:62-90
            cut, fall, use_power, power = t.spot_angles[i]
            origin = tree.vector(cid, "origin", p[0:3])
            direction = tree.vector(cid, "direction", p[3:6])
            c_, f_ = tree.number(cid, "cutoff", cut), tree.number(cid, "falloff", fall)
            if use_power:
                last = f"spot_from_power({tree.color(cid, 'power', power)}, rad({c_}), rad({f_}))"
            else:
                last = tree.color(cid, "intensity", p[8:11])
            s += tree.pull_header() + f"  let light_{cid} = make_spot_light({i}, {origin}, {direction}, rad({c_}), rad({f_}), {last});\n"
        elif ty == LIGHT_SPHERE_AREA:   # AreaLight.cpp:166-190: every entity field is a Dynamic (registry) parameter
            pre = f"ae_{cid}"
            rad = tree.color(cid, "radiance", p[0:3])
            ent = t.entities[int(l["entity_id"])]
            def mat(name, cols):
                return "make_mat%s(%s)" % ("3x4" if len(cols) == 4 else "3x3", ",".join(tree.vector(cid, f"{pre}_{name}_c{k}", c, dynamic=True) for k, c in enumerate(cols)))
            local = mat("local", [ent[3 * k:3 * k + 3] for k in range(4)])
            glob = mat("global", [ent[12 + 3 * k:12 + 3 * k + 3] for k in range(4)])
            norm = mat("normal", [ent[24 + 3 * k:24 + 3 * k + 3] for k in range(3)])
            eid, mid, sid = (tree.integer(cid, f"{pre}_{n}", v) for n, v in (("ent_id", int(l["entity_id"])), ("ent_mat_id", int(ent[34:35].view(np.int32)[0])), ("ent_shp_id", int(ent[33:34].view(np.int32)[0]))))
            org = tree.vector(cid, pre + "_origin", p[3:6], dynamic=True)
            radius = tree.number(cid, pre + "_radius", p[6], dynamic=True)
            s += tree.pull_header() + (f"  let ae_{cid} = make_sphere_area_emitter(Entity{{ id = {eid}, mat_id = {mid}, local_mat = {local}, global_mat = {glob}, "
                                       f"normal_mat = {norm}, shape_id = {sid} }},  Sphere{{ origin = {org}, radius = {radius} }});\n")
            s += f"  let light_{cid} = make_area_light({i}, ae_{cid}, @|ctx| {{ maybe_unused(ctx); {rad} }});\n"
        else:
            raise ValueError(ty)
        fin_names.append(f"light_{cid}")
    s += "\n" if (fin_names or n_embedded) else ""
    for c, count, offset in embedded:
        s += f"  let e_{c.lower()} = load_simple_point_lights({count}, {offset}, device);\n"
        if count == len(t.finite_lights):   # nothing except one embedded class (LoaderLight.cpp:188-193)
            return s + f"  let finite_lights = e_{c.lower()};\n  maybe_unused(finite_lights);\n"
    s += f"  let finite_lights = LightTable {{\n    count = {len(fin_names) + n_embedded},\n    get   = @|id:i32| {{\n"
    for k, (c, count, offset) in enumerate(embedded):
        s += ("    else if " if k else "    if ") + f"id < {offset + count} {{\n      e_{c.lower()}.get(id - {offset})\n    }}\n"
    if embedded:
        s += "    else {\n"
    s += "    match(id) {\n"
    for i, n in enumerate(fin_names):
        s += f"      {i + n_embedded} => {n},\n"
    s += "      _ => make_null_light(id)\n"
    if embedded:
        s += "    }\n"
    s += "    }\n  }};\n  maybe_unused(finite_lights);\n"
    return s


def _technique(t: SceneTables, std_aovs: bool, cache_dir: str | None = None) -> str:
    tech = t.technique
    s = ('  let tech_max_depth = registry::get_global_parameter_i32("__tech_max_depth", 8);\n' if int(tech["max_depth"]) >= 2 else f"  let tech_max_depth = {int(tech['max_depth'])}:i32;\n")
    s += ('  let tech_min_depth = registry::get_global_parameter_i32("__tech_min_depth", 2);\n' if int(tech["min_depth"]) >= 2 else f"  let tech_min_depth = {int(tech['min_depth'])}:i32;\n")
    s += ('  let tech_clamp = registry::get_global_parameter_f32("__tech_clamp", 0);\n' if float(tech["clamp"]) > 0 else f"  let tech_clamp = {_stream(tech['clamp'])}:f32;\n")
    s += "  let aovs = @|id:i32| -> AOVImage {\n    match(id) {\n      _ => make_empty_aov_image(0, 0)\n    }\n  };\n"
    # LoaderLight.cpp:423-452: the cdf ("simple") and hierarchy selectors read a buffer the loader exports into its cache directory
    sel = int(tech["light_selector"]) if "light_selector" in tech.dtype.names else 0
    if sel == 0:
        s += "  let light_selector = make_uniform_light_selector(infinite_lights, finite_lights);\n"
    else:
        if cache_dir is None:
            raise ValueError("the cdf / hierarchy light selectors need a cache directory for their buffer (LoaderLight.cpp:434,441)")
        import os
        name = "light_cdf.bin" if sel == 1 else "light_hierarchy.bin"
        path = os.path.join(cache_dir, name).replace("\\", "/")
        if not os.path.exists(path):
            t.selector_data.astype("<f4").tofile(path)
        if sel == 1:
            s += f'  let light_cdf = cdf::make_cdf_1d_from_buffer(device.load_buffer("{path}"), finite_lights.count, 0);\n'
            s += "  let light_selector = make_cdf_light_selector(infinite_lights, finite_lights, light_cdf);\n"
        else:
            s += f'  let light_selector = make_hierarchy_light_selector(infinite_lights, finite_lights, device.load_buffer("{path}"));\n'
    s += f"  let technique = make_path_renderer(tech_max_depth, tech_min_depth, light_selector, aovs, tech_clamp,{'true' if int(tech['nee']) else 'false'});\n"
    s += ("  let full_technique = wrap_infobuffer_renderer(device, settings.iter, spi, technique);\n" if std_aovs else "  let full_technique = technique;\n")
    return s


def _bsdf(t: SceneTables, mat_id: int, tree: _Tree) -> str:
    m = t.materials[mat_id]
    cid = tree.closure(f"__bsdf_{mat_id}")
    p = m["p"]
    s = ""
    outer = None
    if int(m["map_kind"]) != MAP_NONE:   # MapBSDF.cpp:17-55: the wrapper's closure comes first (map texture, strength), then the inner BSDF's statements
        outer = cid
        tex = tree.texture(t, int(m["map_tex"]))
        strength = tree.number(outer, "strength", m["map_strength"])
        cid = tree.closure(f"__bsdf_{mat_id}_inner")

    def colour(prop, rgb, k, **kw):
        return tree.texture(t, int(m["tex"][k])) if int(m["tex"][k]) >= 0 else tree.color(cid, prop, rgb, **kw)

    if int(m["bsdf"]) == BSDF_DIFFUSE:
        refl = colour("reflectance", p[0:3], 0)
        rough = tree.number(cid, "roughness", 0.0)
        s += tree.pull_header() + f"  let bsdf_{cid} : BSDFShader = @|ctx| make_diffuse_bsdf(ctx.surf, {rough}, {refl});\n"
    elif int(m["bsdf"]) == BSDF_DIELECTRIC:
        ks = colour("specular_reflectance", p[2:5], 0)
        kt = colour("specular_transmittance", p[5:8], 1)
        ext = tree.number(cid, "ext_ior", p[0])
        int_ = tree.number(cid, "int_ior", p[1])
        s += f"  let md_{cid} = @|ctx : ShadingContext| microfacet::make_delta_distribution(ctx.surf.local);\n"
        s += tree.pull_header() + (f"  let bsdf_{cid} : BSDFShader = @|ctx| make_dielectric_bsdf(ctx.surf, {ext}, {int_}, {ks}, {kt}, md_{cid}(ctx), false);\n")
    elif int(m["bsdf"]) == BSDF_CONDUCTOR:   # ConductorBSDF.cpp:13-35 (eta: ColorOptions::Black, k: ColorOptions::White)
        ks = colour("specular_reflectance", p[6:9], 0)
        eta = tree.color(cid, "eta", p[0:3], one=False)    # ColorOptions::Black(): only black is printed into the text
        kk = tree.color(cid, "k", p[3:6], zero=False)       # ColorOptions::White()
        if int(m["distribution"]) == MICROFACET_VNDF_GGX:   # BSDF::setupRoughness, the explicit form (roughness_u / roughness_v; NumberOptions::Zero)
            au, av = tree.number(cid, "roughness_u", m["alpha_u"], one=False), tree.number(cid, "roughness_v", m["alpha_v"], one=False)
            s += tree.pull_header() + (f"  let md_{cid} = @|ctx : ShadingContext| microfacet::make_vndf_ggx_distribution(ctx.surf.face_normal, ctx.surf.local, {au}, {av});\n")
        else:
            s += f"  let md_{cid} = @|ctx : ShadingContext| microfacet::make_delta_distribution(ctx.surf.local);\n"
        s += tree.pull_header() + f"  let bsdf_{cid} : BSDFShader = @|ctx| make_conductor_bsdf(ctx.surf, {eta}, {kk}, {ks}, md_{cid}(ctx));\n"
    else:
        raise ValueError(int(m["bsdf"]))
    if outer is not None:
        if int(m["map_kind"]) == MAP_BUMP:
            tf = f"@|ctx:ShadingContext|->Color{{maybe_unused(ctx); {tex}}}"   # ShadingTree::addTexture, the string case
            s += tree.pull_header() + (f"  let bsdf_{outer} : BSDFShader = @|ctx| make_bumpmap(ctx, @|surf2| -> Bsdf {{  bsdf_{cid}(ctx.{{surf=surf2}}) }}, "
                                       f"texture_dx({tf}, ctx).r, texture_dy({tf}, ctx).r, {strength});\n")
        else:
            s += tree.pull_header() + f"  let bsdf_{outer} : BSDFShader = @|ctx| make_normalmap(ctx, @|surf2| bsdf_{cid}(ctx.{{surf=surf2}}), {tex},{strength});\n"
        cid = outer
    s += "  let medium_interface = no_medium_interface();\n"
    if int(m["light_id"]) >= 0:
        tree.local.ints["_light_id"] = int(m["light_id"])
        s += '  let light_id = registry::get_local_parameter_i32("_light_id", 0);\n'
        s += f"  let shader : MaterialShader = @|ctx| make_emissive_material(mat_id, bsdf_{cid}(ctx), medium_interface, @finite_lights.get(light_id));\n\n"
    else:
        s += f"  let shader : MaterialShader = @|ctx| make_material(mat_id, bsdf_{cid}(ctx), medium_interface);\n\n"
    return s


def generate(t: SceneTables, specialization: str = "default", std_aovs: bool = True, tracer: bool = False, cache_dir: str | None = None) -> StageSet:
    """The stage scripts + registries of one technique variant for `t`. `cache_dir`: where exported buffers go (the loader's cache
    directory, src/runtime/loader/LoaderContext.h CacheManager); only needed by the cdf / hierarchy light selectors."""
    g = Registry()
    cam = t.camera
    g.vectors["__camera_eye"] = tuple(float(x) for x in cam["eye"])
    g.vectors["__camera_dir"] = tuple(float(x) for x in cam["dir"])
    g.vectors["__camera_up"] = tuple(float(x) for x in cam["up"])
    g.vectors["__scene_bbox_lower"] = tuple(float(x) for x in t.bbox_min)
    g.vectors["__scene_bbox_upper"] = tuple(float(x) for x in t.bbox_max)
    g.ints["__entity_count"] = t.n_entities
    g.ints["__material_count"] = int(t.materials.shape[0])
    g.ints["__tech_max_depth"] = int(t.technique["max_depth"])
    g.ints["__tech_min_depth"] = int(t.technique["min_depth"])
    g.floats["__tech_clamp"] = float(t.technique["clamp"])
    prologue, has_sphere = _prologue(t)

    # ---- ray generation
    local = Registry()
    s = STD_LIB_STUB + "#[export] fn ig_ray_generation_shader(settings: &Settings, next_id: i32, size: i32, xmin: i32, ymin: i32, xmax: i32, ymax: i32) -> i32 {\n" + prologue
    s += "  let init_raypayload = make_simple_payload_initializer(init_pt_raypayload);\n"
    if tracer:
        s += "  let emitter = make_list_emitter(device.load_rays(), render_config, init_raypayload);\n"
    else:
        aspect = "(settings.width as f32 / settings.height as f32)" if float(cam["aspect"]) <= 0 else _ts(cam["aspect"])
        fov_gen = "compute_scale_from_vfov" if int(cam["fov_vertical"]) else "compute_scale_from_hfov"
        s += ('  let camera_eye = registry::get_global_parameter_vec3("__camera_eye", vec3_expand(0));\n'
              '  let camera_dir = registry::get_global_parameter_vec3("__camera_dir", vec3_expand(0));\n'
              '  let camera_up  = registry::get_global_parameter_vec3("__camera_up" , vec3_expand(0));\n')
        s += (f"  let camera = make_perspective_camera(camera_eye, \n    camera_dir, \n    camera_up, \n    {fov_gen}({_stream(cam['fov'])}, {aspect}), \n"
              f"    settings.width, \n    settings.height, \n    {_stream(cam['tmin'])}, \n    {_stream(cam['tmax'])});\n\n")
        s += "  let pixel_sampler = make_uniform_pixel_sampler();\n\n"
        s += "  let emitter = make_camera_emitter(camera, render_config, pixel_sampler, init_raypayload);\n"
    s += "  device.generate_rays(emitter, payload_info, GenerateRayInfo{ next_id=next_id, size=size, xmin=xmin, ymin=ymin, xmax=xmax, ymax=ymax })\n}\n"
    raygen = Stage("ig_ray_generation_shader", s, local)

    # ---- miss
    resources: list = []
    local = Registry()
    tree = _Tree(local, specialization)
    tree.resources, tree.cache_dir = resources, cache_dir
    s = STD_LIB_STUB + "#[export] fn ig_miss_shader(settings: &Settings, first: i32, last: i32) -> () {\n" + prologue
    s += _lights(t, tree) + "\n" + _technique(t, std_aovs, cache_dir) + "\n"
    s += "  let use_framebuffer = true;\n  device.handle_miss_shader(full_technique, payload_info, first, last, use_framebuffer);\n}\n"
    miss = Stage("ig_miss_shader", s, local)

    # ---- one hit shader per material
    hits = []
    for mat_id in range(int(t.materials.shape[0])):
        local = Registry()
        tree = _Tree(local, specialization)
        tree.resources, tree.cache_dir = resources, cache_dir
        s = STD_LIB_STUB + "#[export] fn ig_hit_shader(settings: &Settings, mat_id: i32, first: i32, last: i32) -> () {\n" + prologue
        s += _database(has_sphere) + _lights(t, tree) + "\n" + _bsdf(t, mat_id, tree) + _technique(t, std_aovs, cache_dir) + "\n"
        s += "  let use_framebuffer = true;\n  device.handle_hit_shader(shader, scene, full_technique, payload_info, first, last, use_framebuffer);\n}\n"
        hits.append(Stage("ig_hit_shader", s, local))
    fix = {}
    if t.embedded_lights:   # PointLight::embed (PointLight.cpp:70-78): position xyz, pad, intensity rgb, pad
        pts = [l for l in t.finite_lights if int(l["type"]) == LIGHT_POINT]
        tab = np.zeros((len(pts), 8), np.float32)
        for k, l in enumerate(pts):
            tab[k, 0:3], tab[k, 4:7] = l["p"][0:3], l["p"][3:6]
        fix["SimplePointLight"] = tab
    return StageSet(raygen, miss, hits, g, fix, resources)
