// b200_device.cpp -- see b200_device.h. Host-only C++17; every device operation goes through the C ABI.
#include "b200_device.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace igbh {

// ------------------------------------------------------------------------------------------------ compiler device
// The reference JIT-compiles `script` and returns the address of `function` (src/device/Compiler.cpp:13-46,53-64); failure
// is nullptr / false and aborts the load (src/runtime/shader/ShaderManager.cpp:60-64). Here the stage body is parsed into a
// StageDescriptor (script_recognizer.h); the handle is that descriptor. Unrecognised constructs fail the same way.
bool B200CompilerDevice::compile(const Settings& settings, const std::string& script) const {
    // the runtime only uses compile() to warm the JIT cache from `igc` (src/compiler/main.cpp); there is nothing to warm
    (void)settings; (void)script;
    return true;
}

void* B200CompilerDevice::compileAndGet(const Settings& settings, const std::string& script, const std::string& function) const {
    (void)settings;
    try {
        std::unique_ptr<StageDescriptor> d(parse_stage(script, function));
        std::lock_guard<std::mutex> lock(mMutex);
        mStages.push_back(std::move(d));
        return mStages.back().get();
    } catch (const RecognizeError& e) {
        set_last_error(function + ": " + e.what);
        std::fprintf(stderr, "[igb200] cannot compile %s: %s\n", function.c_str(), e.what.c_str());
        return nullptr;
    }
}

// ------------------------------------------------------------------------------------------------ render device
void B200Device::error(const std::string& what) {
    mError = what;
    std::fprintf(stderr, "[igb200] %s\n", what.c_str());   // the reference logs with IG_LOG(L_ERROR) and carries on
}

// One context per GPU. The reference is one Device per process on one GPU (Device.cpp:1632); this device spreads the frame over the GPUs
// named by the environment -- IGB200_GPUS = a count or "all" (default 1: exactly the reference's behaviour) -- starting at the target's device
// index: every context renders its own 32x32 tiles (igb200_comm_init), the frame is assembled on the first GPU when the host or the device
// side asks for it (igb200_comm_gather_framebuffer). Nothing else of the IRenderDevice surface changes, so igcli / igtrace use the GPUs unchanged.
B200Device::B200Device(const SetupSettings& settings) : mSetup(settings) {
    int visible = 0;
    if (igb200_device_count(&visible) != 0) visible = 0;
    int want = 1;
    if (const char* g = std::getenv("IGB200_GPUS")) want = !std::strcmp(g, "all") ? visible : std::atoi(g);
    const int first = (int)settings.target.device();
    if (want < 1 || (want > 1 && first + want > visible)) { error("IGB200_GPUS asks for " + std::to_string(want) + " GPUs from device " + std::to_string(first) + ", " + std::to_string(visible) + " visible"); return; }
    for (int k = 0; k < want; ++k) {
        igb200_ctx* c = nullptr;
        if (igb200_create(first + k, &c) != 0) { error(std::string("cannot create the device: ") + igb200_last_error()); destroyAll(); return; }
        mCtxs.push_back(c);
    }
    mCtx = mCtxs[0];
    if (want > 1) {
        uint8_t id[128];
        if (igb200_comm_unique_id(id) != 0) { error(std::string("NCCL: ") + igb200_last_error()); destroyAll(); return; }
        if (!forAll([&](igb200_ctx* c, int rank) { return igb200_comm_init(c, rank, want, 32, id); }, "igb200_comm_init")) { destroyAll(); return; }
    }
    // BVH cache: the reference's CacheManager lives in the loader (LoaderContext, TriMeshProvider.cpp:326-351) and does not reach the device;
    // this device builds its own trees, so it takes the directory from the environment
    if (const char* dir = std::getenv("IGB200_CACHE_DIR")) { if (*dir) forAll([&](igb200_ctx* c, int) { return igb200_set_cache_dir(c, dir); }, "igb200_set_cache_dir"); }
}

void B200Device::destroyAll() {
    forAll([](igb200_ctx* c, int) { igb200_comm_destroy(c); return 0; }, "igb200_comm_destroy");
    for (igb200_ctx* c : mCtxs) igb200_destroy(c);
    mCtxs.clear(); mCtx = nullptr;
}

B200Device::~B200Device() { destroyAll(); }

bool B200Device::forAll(const std::function<int(igb200_ctx*, int)>& f, const char* what) {
    const int n = (int)mCtxs.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    auto one = [&](int k) { rc[k] = f(mCtxs[k], k); if (rc[k] != 0) msg[k] = igb200_last_error(); };   // the C ABI's last error is per thread
    if (n == 1) one(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < n; ++k) th.emplace_back(one, k);
        for (std::thread& t : th) t.join();
    }
    for (int k = 0; k < n; ++k) if (rc[k] != 0) { error(std::string(what) + " (GPU " + std::to_string(k) + "): " + msg[k]); return false; }
    return true;
}

// Device.cpp:1667-1670: borrowed pointers, valid for the device's lifetime (Runtime.cpp:532-541). The upload itself waits
// for the first render(): the material / light / camera descriptors only arrive with the shader set.
void B200Device::assignScene(const SceneSettings& settings) { mScene = settings; mSceneDirty = true; mFiles = ImageCache(); }

void B200Device::resize(size_t width, size_t height) {
    if (!mCtx) return;
    if (!forAll([&](igb200_ctx* c, int) { return igb200_resize(c, (int)width, (int)height); }, "resize")) return;
    mWidth = width; mHeight = height; mHostPtrs.clear();
}

// Device.cpp:1684-1690 drops cached uploads; the next render() re-uploads the scene.
void B200Device::releaseAll() { mSceneDirty = true; mFiles = ImageCache(); }

bool B200Device::setPartition(int rank, int world, int tile) {
    if (mCtxs.size() > 1) { error("setPartition: this device already spreads the frame over its own GPUs (IGB200_GPUS)"); return false; }
    if (!mCtx || igb200_set_partition(mCtx, rank, world, tile) != 0) { error(igb200_last_error()); return false; }
    return true;
}

static void append(std::vector<uint8_t>& v, const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); v.insert(v.end(), b, b + n); }

// Turns (SceneDatabase, shader set, registries) into igb200_scene_desc and uploads it when anything changed.
bool B200Device::uploadScene(const IG::TechniqueVariantShaderSet& set, const IG::ParameterSet* global) {
    if (!mScene.database || !mScene.entity_per_material) { error("render() before assignScene()"); return false; }
    std::vector<igb200_material> materials;
    std::vector<igb200_light> inf, fin;
    igb200_camera camera; std::memset(&camera, 0, sizeof(camera));
    igb200_technique technique{};
    std::vector<float> selector_data;
    TextureTable textures;   // filled by the materials in order of first use (script_recognizer.h)
    textures.resource_map = mScene.resource_map;
    textures.cache = &mFiles;   // decoded image / buffer files live as long as the scene is assigned: render() resolves the descriptors at every call
    try {
        if (set.HitShaders.size() != mScene.entity_per_material->size()) throw RecognizeError{"one hit shader per material expected"};
        // the stage that carries the light tables: the miss shader, else the first hit shader that has them
        const StageDescriptor* light_stage = nullptr; const IG::ParameterSet* light_local = nullptr;
        for (const auto& hs : set.HitShaders) {
            const StageDescriptor* d = static_cast<const StageDescriptor*>(hs.Exec);
            if (!d) throw RecognizeError{"null hit shader"};
            if (!light_stage && d->has_lights) { light_stage = d; light_local = hs.LocalRegistry.get(); }
        }
        const StageDescriptor* miss = static_cast<const StageDescriptor*>(set.MissShader.Exec);
        if (miss && miss->has_lights) { light_stage = miss; light_local = set.MissShader.LocalRegistry.get(); }
        if (!light_stage) throw RecognizeError{"no stage carries the light tables"};
        // Normals / Albedo AOVs requested by the wrapped technique (runtime flag --no-std-aovs removes the wrapper)
        if ((int)light_stage->std_aovs != mStdAovs) {   // only on a change: setting an option synchronises the device
            if (!forAll([&](igb200_ctx* c, int) { return igb200_set_option(c, "std_aovs", light_stage->std_aovs ? 1 : 0); }, "std_aovs")) throw RecognizeError{mError};
            mStdAovs = (int)light_stage->std_aovs;
        }
        // lights before materials: the textures of environment lights come first in the table, as the loader numbers them
        resolve_lights(*light_stage, Registries{light_local, global}, inf, fin, mScene.database, &textures);
        for (const auto& hs : set.HitShaders)
            materials.push_back(resolve_material(*static_cast<const StageDescriptor*>(hs.Exec), Registries{hs.LocalRegistry.get(), global}, &textures));
        technique = resolve_technique(*light_stage, Registries{light_local, global}, selector_data);
        const StageDescriptor* rg = static_cast<const StageDescriptor*>(set.RayGenerationShader.Exec);
        if (!rg) throw RecognizeError{"null ray generation shader"};
        if (rg->has_camera) camera = resolve_camera(*rg, Registries{set.RayGenerationShader.LocalRegistry.get(), global});
    } catch (const RecognizeError& e) { error("cannot bind the shader set: " + e.what); return false; }

    std::vector<uint8_t> bytes;
    append(bytes, materials.data(), materials.size() * sizeof(igb200_material));
    append(bytes, inf.data(), inf.size() * sizeof(igb200_light));
    append(bytes, fin.data(), fin.size() * sizeof(igb200_light));
    append(bytes, &camera, sizeof(camera));
    append(bytes, &technique, sizeof(technique));
    append(bytes, selector_data.data(), selector_data.size() * sizeof(float));
    append(bytes, textures.records.data(), textures.records.size() * sizeof(igb200_texture));
    for (const std::string& key : textures.image_keys) append(bytes, key.data(), key.size() + 1);   // which files, not their pixels (cached: mFiles)
    append(bytes, textures.aux.data(), textures.aux.size() * sizeof(float));
    if (!mSceneDirty && bytes == mDescriptorBytes) return true;

    const IG::SceneDatabase& db = *mScene.database;
    const auto ent = db.FixTables.find("entities");
    const auto shp = db.DynTables.find("shapes");
    if (ent == db.FixTables.end() || shp == db.DynTables.end()) { error("scene database without 'entities' / 'shapes' tables"); return false; }
    std::vector<uint8_t> leaves;   // EntityLeaf1 of every provider's scene BVH (src/runtime/bvh/SceneBVHAdapter.h:68-100)
    for (const auto& kv : db.SceneBVHs) append(leaves, kv.second.Leaves.data(), kv.second.Leaves.size());
    static_assert(sizeof(igb200_entity_leaf) == 96 && sizeof(igb200_lookup_entry) == sizeof(IG::LookupEntry), "layout");

    igb200_scene_desc d;
    std::memset(&d, 0, sizeof(d));
    d.entities = reinterpret_cast<const float*>(ent->second.data().data());
    d.n_entities = (int32_t)(ent->second.data().size() / (36 * sizeof(float)));
    d.shape_lookups = reinterpret_cast<const igb200_lookup_entry*>(shp->second.lookups().data());
    d.n_shapes = (int32_t)shp->second.lookups().size();
    d.shape_data = shp->second.data().data();
    d.shape_data_bytes = shp->second.data().size();
    d.leaves = reinterpret_cast<const igb200_entity_leaf*>(leaves.data());
    d.n_leaves = (int32_t)(leaves.size() / sizeof(igb200_entity_leaf));
    d.entity_per_material = mScene.entity_per_material->data();
    d.n_materials = (int32_t)materials.size();
    d.materials = materials.data();
    d.infinite_lights = inf.data(); d.n_infinite = (int32_t)inf.size();
    d.finite_lights = fin.data(); d.n_finite = (int32_t)fin.size();
    d.camera = camera; d.technique = technique;
    d.selector_data = selector_data.empty() ? nullptr : selector_data.data(); d.n_selector_data = (int32_t)selector_data.size();
    d.textures = textures.records.empty() ? nullptr : textures.records.data(); d.n_textures = (int32_t)textures.records.size();
    std::vector<igb200_image> images;   // 8-bit files decoded by image_io as the reference's device keeps them
    for (const auto& im : textures.images) { igb200_image ii; ii.format = im->format; ii.width = im->width; ii.height = im->height; ii.reserved = 0; ii.pixels = im->pixels(); images.push_back(ii); }
    d.images = images.empty() ? nullptr : images.data(); d.n_images = (int32_t)images.size();
    d.aux_data = textures.aux.empty() ? nullptr : textures.aux.data(); d.n_aux_data = (int32_t)textures.aux.size();   // 2-D cdfs of textured environment lights
    for (int k = 0; k < 3; ++k) { d.bbox_min[k] = db.SceneBBox.min(k); d.bbox_max[k] = db.SceneBBox.max(k); }
    if (!forAll([&](igb200_ctx* c, int) { return igb200_set_scene(c, &d); }, "scene upload")) return false;   // the scene is replicated
    mDescriptorBytes.swap(bytes);
    mSceneDirty = false;
    return true;
}

// Device.cpp:1672-1682. One iteration, accumulated into the framebuffer. Asynchronous on the device (igb200.h).
void B200Device::render(const IG::TechniqueVariantShaderSet& shader_set, const RenderSettings& settings, IG::ParameterSet* parameter_set) {
    if (!mCtx) { error("render() on an invalid device"); return; }
    if (!uploadScene(shader_set, parameter_set)) return;
    igb200_settings st;
    st.device = (int32_t)mSetup.target.device(); st.thread_count = 0; st.spi = (int32_t)settings.spi; st.frame = (int32_t)settings.frame;
    st.iter = (int32_t)settings.iteration; st.width = (int32_t)settings.width; st.height = (int32_t)settings.height; st.seed = (int32_t)settings.user_seed;
    int rc;
    if (settings.rays) {   // Device.cpp:602-643: directions are normalised on upload
        std::vector<igb200_ray> rays(settings.width);
        for (size_t i = 0; i < settings.width; ++i) {
            const IG::Ray& r = settings.rays[i];
            const float dx = r.Direction(0), dy = r.Direction(1), dz = r.Direction(2);
            const float n = std::sqrt(dx * dx + dy * dy + dz * dz);
            igb200_ray& o = rays[i];
            o.org[0] = r.Origin(0); o.org[1] = r.Origin(1); o.org[2] = r.Origin(2);
            o.dir[0] = dx / n; o.dir[1] = dy / n; o.dir[2] = dz / n;
            o.tmin = r.Range(0); o.tmax = r.Range(1);
        }
        rc = forAll([&](igb200_ctx* c, int) { return igb200_render(c, &st, rays.data(), rays.size()); }, "render") ? 0 : -1;
        mWidth = settings.width; mHeight = 1;
    } else {
        // asynchronous: every GPU is left working on its tiles when this returns
        rc = 0;
        for (igb200_ctx* c : mCtxs) if (igb200_render(c, &st, nullptr, 0) != 0) { rc = -1; error(std::string("render failed: ") + igb200_last_error()); break; }
        mWidth = settings.width; mHeight = settings.height;
    }
    (void)rc;
}

static const char* aov_name(const std::string& name) { return (name.empty() || name == "Color") ? nullptr : name.c_str(); }

// Device.cpp:1419-1451: RGB f32, W*H*3, owned by the device, valid until resize.
IG::IRenderDevice::AOVAccessor B200Device::getFramebufferForHost(const std::string& name, bool) {
    float* p = nullptr;
    if (!mCtx) { error("no device"); return AOVAccessor{nullptr}; }
    if (mCtxs.size() > 1) {   // every GPU sends its tiles to the first one, which assembles the frame and copies it to pinned host memory
        const char* aov = aov_name(name) ? name.c_str() : "";
        if (!forAll([&](igb200_ctx* c, int rank) { return igb200_comm_gather_framebuffer(c, aov, nullptr, rank == 0 ? &p : nullptr); }, "framebuffer gather")) return AOVAccessor{nullptr};
    } else if (igb200_framebuffer(mCtx, aov_name(name), &p) != 0) { error(igb200_last_error()); return AOVAccessor{nullptr}; }
    mHostPtrs[aov_name(name) ? name : std::string()] = p;
    return AOVAccessor{p};
}
IG::IRenderDevice::AOVAccessor B200Device::getFramebufferForDevice(const std::string& name, bool) {
    float* p = nullptr;
    if (!mCtx) { error("no device"); return AOVAccessor{nullptr}; }
    if (mCtxs.size() > 1) {
        const char* aov = aov_name(name) ? name.c_str() : "";
        if (!forAll([&](igb200_ctx* c, int rank) { return igb200_comm_gather_framebuffer(c, aov, rank == 0 ? &p : nullptr, nullptr); }, "framebuffer gather")) return AOVAccessor{nullptr};
        if (igb200_sync(mCtx) != 0) { error(igb200_last_error()); return AOVAccessor{nullptr}; }   // the assembled frame is complete when the pointer is handed out
    } else if (igb200_framebuffer_device(mCtx, aov_name(name), &p) != 0) { error(igb200_last_error()); return AOVAccessor{nullptr}; }
    return AOVAccessor{p};
}
void B200Device::clearFramebuffer(const std::string& name) { if (mCtx) forAll([&](igb200_ctx* c, int) { return igb200_clear(c, aov_name(name)); }, "clear"); }
void B200Device::clearAllFramebuffer() { if (mCtx) forAll([](igb200_ctx* c, int) { return igb200_clear(c, nullptr); }, "clear"); }
// Device.cpp:1724-1736: pushes the host copy the caller may have modified back to the device
void B200Device::syncFramebufferHostToDevice(const std::string& name) {
    if (!mCtx) return;
    const auto it = mHostPtrs.find(aov_name(name) ? name : std::string());   // the host copy of THIS AOV, never another one's
    if (it == mHostPtrs.end() || !it->second) return;                      // never mapped for the host: nothing the caller could have changed
    // several GPUs: each keeps the whole frame but only its own tiles are ever gathered, so all of them take the upload
    float* host = it->second;
    forAll([&](igb200_ctx* c, int) { return igb200_upload_framebuffer(c, aov_name(name), host); }, "framebuffer upload");
}
// Device.cpp:1724-1736 maps every AOV: here every AOV the host has a copy of
void B200Device::syncAllFramebufferHostToDevice() {
    std::vector<std::string> names;
    for (const auto& kv : mHostPtrs) names.push_back(kv.first);
    for (const std::string& n : names) syncFramebufferHostToDevice(n);
}

// Named scratch buffers belong to techniques outside this path (photon mapper, AEPT): none exist here.
size_t B200Device::getBufferSizeInBytes(const std::string&) { return 0; }
bool B200Device::copyBufferToHost(const std::string& name, void*, size_t) { error("buffer '" + name + "' does not exist on this device"); return false; }
IG::IRenderDevice::BufferAccessor B200Device::getBufferForDevice(const std::string&) { return BufferAccessor{nullptr, 0}; }

// Device.cpp:1762-1771. The reference's Statistics keeps its counters private (the runtime merges and prints them), so the device
// refills the object from the device-side totals; rayCounters() is the same figures for callers without the class.
const IG::Statistics* B200Device::getStatistics() {
    uint64_t s[3]; double ms = 0;
    if (!rayCounters(s, &ms)) return nullptr;
    mStats.reset();
    mStats.increase(IG::Quantity::CameraRayCount, s[0]); mStats.increase(IG::Quantity::ShadowRayCount, s[1]); mStats.increase(IG::Quantity::BounceRayCount, s[2]);
    return &mStats;
}
bool B200Device::rayCounters(uint64_t out[3], double* render_ms) {
    if (!mCtx) return false;
    out[0] = out[1] = out[2] = 0;
    double ms_max = 0;
    for (igb200_ctx* c : mCtxs) {   // rays add up over the GPUs, time is the slowest GPU's
        uint64_t s[5]; double ms = 0;
        if (igb200_stats(c, s, &ms) != 0) return false;
        out[0] += s[0]; out[1] += s[1]; out[2] += s[2];
        if (ms > ms_max) ms_max = ms;
    }
    if (render_ms) *render_ms = ms_max;
    return true;
}

// Post kernels are outside the hot path: reported, never silently emulated.
void B200Device::tonemap(uint32_t*, const IG::TonemapSettings&) { error("tonemap is not supported by the B200 device (viewer-only post kernel)"); }
IG::ImageInfoOutput B200Device::imageinfo(const IG::ImageInfoSettings&) { error("imageinfo is not supported by the B200 device"); return IG::ImageInfoOutput{}; }
void B200Device::bake(const IG::ShaderOutput<void*>&, const std::vector<std::string>*, float*) { error("bake is not supported by the B200 device"); }
void B200Device::runPass(const IG::ShaderOutput<void*>&) { error("runPass is not supported by the B200 device"); }

IG::IRenderDevice* B200DeviceInterface::createRenderDevice(const IG::IRenderDevice::SetupSettings& settings) const {
    B200Device* d = new B200Device(settings);
    if (!d->valid()) { delete d; return nullptr; }
    return d;
}

}  // namespace igbh

// src/device/Interface.cpp:70-76
extern "C" const IG::IDeviceInterface* ig_get_interface() {
    static const igbh::B200DeviceInterface interface;
    return &interface;
}
