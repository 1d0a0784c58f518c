// script_recognizer.h -- the B200 device's replacement for the reference's JIT compiler device.
//
// The reference hands its device plugin *program text*: `ICompilerDevice::compileAndGet(settings, script, function)`
// (src/runtime/device/ICompilerDevice.h:15, called from src/runtime/Runtime.cpp:631-657) receives the whole Artic
// standard library followed by ONE generated stage function (`ig_hit_shader`, `ig_ray_generation_shader`, ...), and the
// AnyDSL plugin JIT-compiles it (src/device/Compiler.cpp:13-46). A native device has no Artic compiler. What it needs
// from the text is small and regular: the generators (src/runtime/shader/*.cpp, src/runtime/bsdf/*.cpp,
// src/runtime/light/*.cpp, src/runtime/camera/PerspectiveCamera.cpp, src/runtime/technique/PathTechnique.cpp) only
// ever emit `let` bindings whose right-hand sides are constructor calls with literal or registry-lookup arguments
// (SURVEY.md Appendix D). This module parses the stage body into those bindings and evaluates the constructor
// arguments against the stage's LocalRegistry and the global registry, yielding the plain descriptors of
// include/igb200.h. Anything it does not recognise is an error (never a silent default): the stage handle is null
// and the reason is in last_error(), which is how the reference reports a failed compile (ShaderManager.cpp:60-64).
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/igb200.h"
#include "ig_mirror.h"
#include "image_io.h"

namespace igbh {

enum class StageKind { RayGeneration, Miss, Hit, Traversal, Device, Callback, AdvancedShadow, Tonemap, ImageInfo, Pass, Bake };

// One `let name [: type] = expr;` of the stage body, expression kept as text until the registries are known.
struct Binding { std::string name, expr; };

// Opaque handle returned by compileAndGet: the parsed stage. Owned by the compiler device's cache.
struct StageDescriptor {
    StageKind kind;
    std::string function;
    std::vector<Binding> lets;                 // in source order
    std::map<std::string, size_t> index;       // name -> position in lets (last definition wins)
    // structure found at parse time (names of bindings; values are resolved later)
    std::string bsdf_binding;                  // hit: `bsdf_<id>` used by the material shader
    bool emissive = false;                     // hit: make_emissive_material(..., @finite_lights.get(light_id))
    std::vector<std::string> infinite_lights;  // hit / miss: `light_<id>` bindings in table order
    std::vector<std::string> finite_lights;    // ... or "@<EmbedClass>:<k>": entry k of the embedded fix-table of that class (LoaderLight.cpp:171-236)
    bool has_lights = false, has_technique = false, has_camera = false, list_emitter = false;
    bool std_aovs = false;                     // the technique is wrapped: wrap_infobuffer_renderer (Normals / Albedo AOVs, InfoBufferTechnique.cpp:13-17)
};

// Thrown by the evaluator; caught at the API boundary and turned into a null handle / false + last_error().
struct RecognizeError { std::string what; };

// ---- parse ---------------------------------------------------------------------------------------------------------
// Finds `fn <function>(` in `script` (the stage function is the last one in the text) and parses its body.
StageDescriptor* parse_stage(const std::string& script, const std::string& function);   // throws RecognizeError

// ---- resolve against registries ------------------------------------------------------------------------------------
struct Registries { const IG::ParameterSet* local; const IG::ParameterSet* global; };

// The scene's texture table as it builds up while the materials are resolved (hit stages in material order): a texture gets the index of
// its first use, as in the loader (ignis_b200/scene.py texture_id); the same texture met again in another stage (its `tex_<id>` binding may
// carry another closure id there) is found by its contents.
// Decoded files kept across the render() calls of a device (the descriptors are resolved at every call to notice changed parameters; the files
// behind resource ids do not change while a scene is assigned): key = path + how it was read.
struct ImageCache {
    std::map<std::string, std::shared_ptr<const DeviceImage>> images;
    std::map<std::string, std::shared_ptr<const std::vector<float>>> buffers;
};

struct TextureTable {
    std::vector<igb200_texture> records;
    int add(const igb200_texture& t);
    // image textures: the files named by the stage text through resource ids (IRenderDevice::SceneSettings::resource_map), decoded as the
    // reference's device keeps them (image_io.h); one entry per (file, linear flag), numbered by first use
    const std::vector<std::string>* resource_map = nullptr;
    ImageCache* cache = nullptr;                        // optional: decoded files are looked up / kept there
    std::vector<std::shared_ptr<const DeviceImage>> images;
    std::vector<std::string> image_keys;                // "<path>|srgb", "<path>|linear", "<path>|float"
    int image(const std::string& path, bool linear);   // 8-bit file (PNG); throws RecognizeError
    int float_image(const std::string& path);          // OpenEXR file
    std::shared_ptr<const std::vector<float>> buffer(const std::string& path);   // raw 32-bit words of a buffer file
    // 32-bit words of further buffers the descriptors point into (igb200_scene_desc::aux_data): the 2-D cdfs of textured environment lights
    std::vector<float> aux;
};
// textures: null = the stage must not use any (an error otherwise)
igb200_material resolve_material(const StageDescriptor& hit, const Registries& r, TextureTable* textures = nullptr);     // throws RecognizeError
// db: the scene database, needed when the finite lights come from embedded fix-tables (load_simple_point_lights reads FixTables["SimplePointLight"],
// LoaderLight.cpp:171-236,397-422; light/point.art:20-36); null = such tables are an error
// textures: needed when an environment light is textured (environment maps, the sky): its texture, image and 2-D cdf (the buffer file named through
// the resource map) are entered there; null = such lights are an error. Resolve the lights BEFORE the materials to number textures as the loader does.
void resolve_lights(const StageDescriptor& stage, const Registries& r, std::vector<igb200_light>& infinite, std::vector<igb200_light>& finite, const IG::SceneDatabase* db = nullptr,
                    TextureTable* textures = nullptr);
// selector_data: contents of the buffer the cdf / hierarchy light selector reads (igb200_scene_desc::selector_data), empty for uniform
igb200_technique resolve_technique(const StageDescriptor& stage, const Registries& r, std::vector<float>& selector_data);
igb200_camera resolve_camera(const StageDescriptor& raygen, const Registries& r);

const std::string& last_error();
void set_last_error(const std::string& e);

}  // namespace igbh
