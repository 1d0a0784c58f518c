// host_capi.cpp -- C shim over the C++ plugin classes so that the Python tests can play the part of IG::Runtime:
// build a SceneDatabase and registries, "compile" stage scripts through ICompilerDevice, drive IRenderDevice.
// Everything goes through ig_get_interface(), i.e. exactly the path DeviceManager takes (src/runtime/device/DeviceManager.cpp:196-218).
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "b200_device.h"

#include <algorithm>

using namespace IG;

namespace {
struct ShaderSet { TechniqueVariantShaderSet set; };
std::shared_ptr<ParameterSet> share(ParameterSet* p) { return p ? std::shared_ptr<ParameterSet>(p, [](ParameterSet*) {}) : nullptr; }
ICompilerDevice* compiler() { static std::unique_ptr<ICompilerDevice> c(ig_get_interface()->createCompilerDevice()); return c.get(); }
}  // namespace

extern "C" {

const char* igbh_last_error() { return igbh::last_error().c_str(); }
int igbh_interface_version(int* major, int* minor) { const Build::Version v = ig_get_interface()->getVersion(); *major = (int)v.Major; *minor = (int)v.Minor; return ig_get_interface()->getArchitecture() == TargetArchitecture{GPUArchitecture::Nvidia} ? 0 : -1; }

// ---- SceneDatabase
SceneDatabase* igbh_db_create() { return new SceneDatabase(); }
void igbh_db_destroy(SceneDatabase* db) { delete db; }
void igbh_db_set_fix(SceneDatabase* db, const char* name, const uint8_t* bytes, size_t n, size_t entries) {
    FixTable t;
    for (size_t i = 0; i < entries; ++i) { auto& d = t.addEntry(0); const size_t per = n / entries; d.insert(d.end(), bytes + i * per, bytes + (i + 1) * per); }
    db->FixTables[name] = std::move(t);
}
void igbh_db_set_dyn(SceneDatabase* db, const char* name, const igb200_lookup_entry* lookups, size_t n_lookups, const uint8_t* data, size_t n) {
    DynTable t;
    for (size_t i = 0; i < n_lookups; ++i) {
        const size_t b = (size_t)lookups[i].offset, e = i + 1 < n_lookups ? (size_t)lookups[i + 1].offset : n;
        auto& d = t.addLookup(lookups[i].type_id, lookups[i].flags, 0);
        d.insert(d.end(), data + b, data + e);
    }
    db->DynTables[name] = std::move(t);
}
void igbh_db_set_bvh(SceneDatabase* db, int provider, const uint8_t* leaves, size_t n) {
    static const char* names[] = {"trimesh", "sphere"};   // ShapeProvider::identifier(), TriMeshProvider.h:12, SphereProvider.h:11
    SceneBVH b; b.Leaves.assign(leaves, leaves + n);
    db->SceneBVHs[std::string_view(names[provider ? 1 : 0])] = std::move(b);
}
void igbh_db_set_bbox(SceneDatabase* db, const float mn[3], const float mx[3], size_t materials) {
    for (int k = 0; k < 3; ++k) { db->SceneBBox.min(k) = mn[k]; db->SceneBBox.max(k) = mx[k]; }
    db->MaterialCount = materials;
}

// ---- ParameterSet
ParameterSet* igbh_params_create() { return new ParameterSet(); }
void igbh_params_destroy(ParameterSet* p) { delete p; }
void igbh_params_set_int(ParameterSet* p, const char* k, int v) { p->IntParameters[k] = v; }
void igbh_params_set_float(ParameterSet* p, const char* k, float v) { p->FloatParameters[k] = v; }
void igbh_params_set_vec3(ParameterSet* p, const char* k, const float v[3]) { p->VectorParameters[k] = Vector3f(v[0], v[1], v[2]); }
void igbh_params_set_color(ParameterSet* p, const char* k, const float v[4]) { p->ColorParameters[k] = Vector4f(v[0], v[1], v[2], v[3]); }

// ---- ICompilerDevice
void* igbh_compile(const char* script, const char* function) { return compiler()->compileAndGet(ICompilerDevice::Settings{}, script, function); }

// descriptors of a compiled stage resolved against registries (recogniser tests, no GPU needed)
int igbh_describe_material(void* stage, ParameterSet* local, ParameterSet* global, igb200_material* out) {
    try { *out = igbh::resolve_material(*static_cast<igbh::StageDescriptor*>(stage), igbh::Registries{local, global}); return 0; }
    catch (const igbh::RecognizeError& e) { igbh::set_last_error(e.what); return -1; }
}
// ... with a texture table that persists over the hit stages of a scene (created / destroyed by the caller)
struct TexHandle : igbh::TextureTable { std::vector<std::string> res; };
igbh::TextureTable* igbh_textures_create() { return new TexHandle(); }
void igbh_textures_destroy(igbh::TextureTable* t) { delete static_cast<TexHandle*>(t); }
void igbh_textures_set_resources(igbh::TextureTable* t, const char* const* paths, int n) {   // SceneSettings::resource_map for the recogniser tests
    TexHandle* h = static_cast<TexHandle*>(t);
    h->res.assign(paths, paths + n);
    h->resource_map = &h->res;
}
int igbh_textures_image_count(const igbh::TextureTable* t) { return (int)t->images.size(); }
const uint8_t* igbh_textures_image(const igbh::TextureTable* t, int i, int* format, int* width, int* height, size_t* bytes) {
    const igbh::DeviceImage& im = *t->images[(size_t)i];
    *format = im.format; *width = im.width; *height = im.height; *bytes = im.pixel_bytes();
    return static_cast<const uint8_t*>(im.pixels());
}
const float* igbh_textures_aux(const igbh::TextureTable* t, size_t* words) { *words = t->aux.size(); return t->aux.data(); }
const uint8_t* igbh_srgb_lut() { return igbh::srgb_byte_to_linear_byte(); }
// test hook: an OpenEXR file as the device keeps it (image_io.h load_float_image); returns the pixel count * 4 or -1; out may be null (size query)
long igbh_load_float_image(const char* path, int* width, int* height, float* out, long cap) {
    try {
        const igbh::FloatImage im = igbh::load_float_image(path);
        *width = im.width; *height = im.height;
        if (out && (long)im.rgba.size() <= cap) std::memcpy(out, im.rgba.data(), im.rgba.size() * sizeof(float));
        return (long)im.rgba.size();
    } catch (const igbh::RecognizeError& e) { igbh::set_last_error(e.what); return -1; }
}
int igbh_textures_count(const igbh::TextureTable* t) { return (int)t->records.size(); }
void igbh_textures_get(const igbh::TextureTable* t, int i, igb200_texture* out) { *out = t->records[(size_t)i]; }
int igbh_describe_material_tex(void* stage, ParameterSet* local, ParameterSet* global, igbh::TextureTable* textures, igb200_material* out) {
    try { *out = igbh::resolve_material(*static_cast<igbh::StageDescriptor*>(stage), igbh::Registries{local, global}, textures); return 0; }
    catch (const igbh::RecognizeError& e) { igbh::set_last_error(e.what); return -1; }
}
int igbh_describe_lights_db(void* stage, ParameterSet* local, ParameterSet* global, const SceneDatabase* db, igb200_light* inf, int* n_inf, igb200_light* fin, int* n_fin, int cap);
int igbh_describe_lights(void* stage, ParameterSet* local, ParameterSet* global, igb200_light* inf, int* n_inf, igb200_light* fin, int* n_fin, int cap) {
    return igbh_describe_lights_db(stage, local, global, nullptr, inf, n_inf, fin, n_fin, cap);
}
int igbh_describe_lights_tex(void* stage, ParameterSet* local, ParameterSet* global, const SceneDatabase* db, igbh::TextureTable* textures, igb200_light* inf, int* n_inf, igb200_light* fin, int* n_fin, int cap);
int igbh_describe_lights_db(void* stage, ParameterSet* local, ParameterSet* global, const SceneDatabase* db, igb200_light* inf, int* n_inf, igb200_light* fin, int* n_fin, int cap) {
    return igbh_describe_lights_tex(stage, local, global, db, nullptr, inf, n_inf, fin, n_fin, cap);
}
int igbh_describe_lights_tex(void* stage, ParameterSet* local, ParameterSet* global, const SceneDatabase* db, igbh::TextureTable* textures, igb200_light* inf, int* n_inf, igb200_light* fin, int* n_fin, int cap) {
    try {
        std::vector<igb200_light> a, b;
        igbh::resolve_lights(*static_cast<igbh::StageDescriptor*>(stage), igbh::Registries{local, global}, a, b, db, textures);
        if ((int)a.size() > cap || (int)b.size() > cap) { igbh::set_last_error("too many lights for the output arrays"); return -1; }
        std::memcpy(inf, a.data(), a.size() * sizeof(igb200_light)); std::memcpy(fin, b.data(), b.size() * sizeof(igb200_light));
        *n_inf = (int)a.size(); *n_fin = (int)b.size();
        return 0;
    } catch (const igbh::RecognizeError& e) { igbh::set_last_error(e.what); return -1; }
}
// selector: receives up to `cap` words of the light selector's buffer; *n = the number of words it holds (may exceed cap)
int igbh_describe_technique(void* stage, ParameterSet* local, ParameterSet* global, igb200_technique* out, float* selector, int cap, int* n) {
    try {
        std::vector<float> data;
        *out = igbh::resolve_technique(*static_cast<igbh::StageDescriptor*>(stage), igbh::Registries{local, global}, data);
        if (n) *n = (int)data.size();
        if (selector) std::copy(data.begin(), data.begin() + std::min<size_t>(data.size(), (size_t)std::max(cap, 0)), selector);
        return 0;
    }
    catch (const igbh::RecognizeError& e) { igbh::set_last_error(e.what); return -1; }
}
int igbh_describe_camera(void* stage, ParameterSet* local, ParameterSet* global, igb200_camera* out) {
    try { *out = igbh::resolve_camera(*static_cast<igbh::StageDescriptor*>(stage), igbh::Registries{local, global}); return 0; }
    catch (const igbh::RecognizeError& e) { igbh::set_last_error(e.what); return -1; }
}

// ---- shader set
ShaderSet* igbh_set_create() { return new ShaderSet(); }
void igbh_set_destroy(ShaderSet* s) { delete s; }
void igbh_set_raygen(ShaderSet* s, void* stage, ParameterSet* local) { s->set.RayGenerationShader = ShaderOutput<void*>{stage, share(local)}; }
void igbh_set_miss(ShaderSet* s, void* stage, ParameterSet* local) { s->set.MissShader = ShaderOutput<void*>{stage, share(local)}; }
void igbh_set_add_hit(ShaderSet* s, void* stage, ParameterSet* local) { s->set.HitShaders.push_back(ShaderOutput<void*>{stage, share(local)}); }

// ---- IRenderDevice
IRenderDevice* igbh_device_create(int cuda_device) {
    IRenderDevice::SetupSettings st; st.target = Target::makeGPU(GPUArchitecture::Nvidia, (size_t)cuda_device);
    IRenderDevice* d = ig_get_interface()->createRenderDevice(st);
    if (!d) igbh::set_last_error(igb200_last_error());
    return d;
}
void igbh_device_destroy(IRenderDevice* d) { delete d; }
struct AssignKeep { std::vector<int32> epm; std::vector<std::string> resources; };
void* igbh_device_assign_res(IRenderDevice* d, SceneDatabase* db, const int32_t* entity_per_material, size_t n, const char* const* resources, int n_resources) {
    AssignKeep* k = new AssignKeep{std::vector<int32>(entity_per_material, entity_per_material + n), {}};   // borrowed by the device (Runtime.cpp:532-541)
    if (resources) k->resources.assign(resources, resources + n_resources);
    IRenderDevice::SceneSettings s; s.database = db; s.entity_per_material = &k->epm; s.resource_map = &k->resources;
    d->assignScene(s);
    return k;
}
void* igbh_device_assign(IRenderDevice* d, SceneDatabase* db, const int32_t* entity_per_material, size_t n) { return igbh_device_assign_res(d, db, entity_per_material, n, nullptr, 0); }
void igbh_assign_release(void* keep) { delete static_cast<AssignKeep*>(keep); }
int igbh_device_render(IRenderDevice* d, ShaderSet* s, ParameterSet* global, int spi, int width, int height, int iteration, int frame, int seed, const float* rays, size_t n_rays) {
    IRenderDevice::RenderSettings rs;
    rs.spi = (size_t)spi; rs.width = (size_t)width; rs.height = (size_t)height; rs.iteration = (size_t)iteration; rs.frame = (size_t)frame; rs.user_seed = (size_t)seed;
    std::vector<Ray> r(n_rays);
    if (rays) {
        for (size_t i = 0; i < n_rays; ++i) { const float* p = rays + 8 * i; r[i] = Ray{Vector3f(p[0], p[1], p[2]), Vector3f(p[3], p[4], p[5]), Vector2f(p[6], p[7])}; }
        rs.rays = r.data(); rs.width = n_rays; rs.height = 1;
    }
    igbh::B200Device* bd = static_cast<igbh::B200Device*>(d);
    const std::string before = bd->lastError();
    d->render(s->set, rs, global);
    if (bd->lastError() != before) { igbh::set_last_error(bd->lastError()); return -1; }
    return 0;
}
void igbh_device_resize(IRenderDevice* d, int w, int h) { d->resize((size_t)w, (size_t)h); }
float* igbh_device_framebuffer(IRenderDevice* d, const char* name) {
    float* p = d->getFramebufferForHost(name ? name : "").Data;
    if (!p) igbh::set_last_error(static_cast<igbh::B200Device*>(d)->lastError());
    return p;
}
void igbh_device_clear(IRenderDevice* d) { d->clearAllFramebuffer(); }
int igbh_device_gpu_count(IRenderDevice* d) { return static_cast<igbh::B200Device*>(d)->gpuCount(); }
int igbh_device_stats(IRenderDevice* d, uint64_t out[3]) {
    if (!d->getStatistics()) return -1;   // the interface call (its counters are private, as in the reference) ...
    return static_cast<igbh::B200Device*>(d)->rayCounters(out) ? 0 : -1;   // ... and the figures it was filled from
}

}  // extern "C"
