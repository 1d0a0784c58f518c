// ig_mirror.h -- the part of the reference's device-plugin contract that the B200 host layer implements, restated.
//
// The reference's plugin boundary is a C++ vtable: `ig_get_interface()` returns an `IDeviceInterface`
// (src/runtime/device/IDeviceInterface.h:9-17) that creates an `IRenderDevice` (src/runtime/device/IRenderDevice.h:14-81)
// and an `ICompilerDevice` (src/runtime/device/ICompilerDevice.h:6-16). Its real headers drag in Eigen, TBB and the
// generated `generated_interface.h`, none of which exist in this container, so this file restates ONLY the types that
// cross the boundary on the `path` hot path, with the reference's names, member names and meaning. Inside the reference's
// build tree, define IGB200_WITH_IGNIS and the real headers are used instead; `namespace IG` is the
// same in both cases, so b200_device.cpp compiles unchanged.
#pragma once

#ifdef IGB200_WITH_IGNIS
#include "device/IDeviceInterface.h"
#include "table/SceneDatabase.h"
#else

#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace IG {

using uint8 = uint8_t;
using int32 = int32_t;
using uint32 = uint32_t;
using uint64 = uint64_t;

// Stand-ins for the Eigen vectors of the reference (same storage: packed floats)
struct Vector2f { float v[2]; float x() const { return v[0]; } float y() const { return v[1]; } };
struct Vector3f { float v[3]; float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; } };
struct Vector4f { float v[4]; float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; } float w() const { return v[3]; } };

// src/runtime/RuntimeStructs.h:39-43
struct Ray { Vector3f Origin; Vector3f Direction; Vector2f Range; };

// src/runtime/ParameterSet.h:6-12 -- the global registry and the per-stage local registries
struct ParameterSet {
    std::unordered_map<std::string, int> IntParameters;
    std::unordered_map<std::string, float> FloatParameters;
    std::unordered_map<std::string, Vector3f> VectorParameters;
    std::unordered_map<std::string, Vector4f> ColorParameters;
    std::unordered_map<std::string, std::string> StringParameters;
};

// src/runtime/table/{FixTable.h:9-31, DynTable.h:6-36, SceneDatabase.h:8-21}
struct LookupEntry { uint32 TypeID; uint32 Flags; uint64 Offset; };
class FixTable {
public:
    std::vector<uint8>& addEntry(size_t alignment) { pad(alignment); ++mCount; return mData; }
    const std::vector<uint8>& data() const { return mData; }
    size_t entryCount() const { return mCount; }
private:
    void pad(size_t a) { if (a != 0 && !mData.empty()) mData.resize(mData.size() + (a - mData.size() % a)); }   // as FixTable.h:17-20
    size_t mCount = 0;
    std::vector<uint8> mData;
};
class DynTable {
public:
    std::vector<uint8>& addLookup(uint32 typeID, uint32 flags, size_t alignment) {
        if (alignment != 0 && !mData.empty()) mData.resize(mData.size() + (alignment - mData.size() % alignment));
        mLookups.push_back(LookupEntry{typeID, flags, (uint64)mData.size()});
        return mData;
    }
    size_t entryCount() const { return mLookups.size(); }
    const std::vector<LookupEntry>& lookups() const { return mLookups; }
    const std::vector<uint8>& data() const { return mData; }
private:
    std::vector<LookupEntry> mLookups;
    std::vector<uint8> mData;
};
struct BoundingBox { Vector3f min, max; };
struct SceneBVH { std::vector<uint8> Nodes; std::vector<uint8> Leaves; };
struct SceneDatabase {
    std::unordered_map<std::string_view, SceneBVH> SceneBVHs;   // key = shape provider ("trimesh", "sphere")
    std::unordered_map<std::string, DynTable> DynTables;        // "shapes"
    std::unordered_map<std::string, FixTable> FixTables;        // "entities"
    float SceneRadius = 0;
    BoundingBox SceneBBox{};
    size_t MaterialCount = 0;
};

// src/runtime/device/Target.h:7-21 (only what a plugin answers with)
enum class GPUArchitecture { AMD_HSA, Intel, Nvidia, Unknown };
struct Target {
    bool gpu = true; GPUArchitecture arch = GPUArchitecture::Nvidia; size_t dev = 0;
    bool isGPU() const { return gpu; }
    GPUArchitecture gpuArchitecture() const { return arch; }
    size_t device() const { return dev; }
};

// src/runtime/technique/TechniqueVariant.h:5-35
enum class CallbackType { BeforeIteration = 0, AfterIteration, _COUNT };
template <typename T> struct ShaderOutput { T Exec; std::shared_ptr<ParameterSet> LocalRegistry; };
template <typename T> struct TechniqueVariantBase {
    uint32 ID = 0;
    ShaderOutput<T> DeviceShader, TonemapShader, ImageinfoShader, PrimaryTraversalShader, SecondaryTraversalShader, RayGenerationShader, MissShader;
    std::vector<ShaderOutput<T>> HitShaders, AdvancedShadowHitShaders, AdvancedShadowMissShaders;
    std::array<ShaderOutput<T>, (size_t)CallbackType::_COUNT> CallbackShaders{};
};
using TechniqueVariantShaderSet = TechniqueVariantBase<void*>;
struct TechniqueVariantInfo { bool UsesLights = true; size_t PrimaryPayloadCount = 6, SecondaryPayloadCount = 0; };   // technique/TechniqueInfo.h (fields used here)

// src/runtime/Statistics.h:57-64 (the ray counters; the rest of the class is host-side bookkeeping)
class Statistics {
public:
    uint64 CameraRayCount = 0, ShadowRayCount = 0, BounceRayCount = 0;
    double RenderMilliseconds = 0;
    uint64 primaryRays() const { return CameraRayCount + BounceRayCount; }               // Statistics.cpp:286-290
    uint64 totalRays() const { return CameraRayCount + BounceRayCount + ShadowRayCount; }
};

struct TonemapSettings; struct ImageInfoSettings;
struct ImageInfoOutput { float Min = 0, Max = 0, Average = 0; };

namespace Build { struct Version { uint32 Major, Minor; uint32 asNumber() const { return (Major << 16) | Minor; } }; }

// src/runtime/device/IRenderDevice.h:14-81
class IRenderDevice {
public:
    struct SetupSettings { Target target; bool AcquireStats = false; bool DebugTrace = false; bool IsInteractive = false; };
    struct SceneSettings {
        SceneDatabase* database = nullptr;
        const std::vector<std::string>* aov_map = nullptr;
        const std::vector<std::string>* resource_map = nullptr;
        const std::vector<int32>* entity_per_material = nullptr;
    };
    struct RenderSettings {
        const Ray* rays = nullptr;   // non-null: width = number of rays, height = 1
        size_t spi = 8, width = 0, height = 0, iteration = 0, frame = 0, user_seed = 0;
        TechniqueVariantInfo info;
    };
    struct AOVAccessor { float* Data; };
    struct BufferAccessor { void* Data; size_t SizeInBytes; };

    virtual ~IRenderDevice() = default;
    virtual void assignScene(const SceneSettings& settings) = 0;
    virtual void render(const TechniqueVariantShaderSet& shader_set, const RenderSettings& settings, ParameterSet* parameter_set) = 0;
    virtual void resize(size_t width, size_t height) = 0;
    virtual void releaseAll() = 0;
    virtual Target target() const = 0;
    virtual size_t framebufferWidth() const = 0;
    virtual size_t framebufferHeight() const = 0;
    virtual bool isInteractive() const = 0;
    virtual AOVAccessor getFramebufferForHost(const std::string& name, bool sync = true) = 0;
    virtual AOVAccessor getFramebufferForDevice(const std::string& name, bool sync = true) = 0;
    virtual void clearFramebuffer(const std::string& name) = 0;
    virtual void clearAllFramebuffer() = 0;
    virtual void syncFramebufferHostToDevice(const std::string& name) = 0;
    virtual void syncAllFramebufferHostToDevice() = 0;
    virtual size_t getBufferSizeInBytes(const std::string& name) = 0;
    virtual bool copyBufferToHost(const std::string& name, void* buffer, size_t maxSizeByte) = 0;
    virtual BufferAccessor getBufferForDevice(const std::string& name) = 0;
    virtual const Statistics* getStatistics() = 0;
    virtual void tonemap(uint32_t*, const TonemapSettings&) = 0;
    virtual ImageInfoOutput imageinfo(const ImageInfoSettings&) = 0;
    virtual void bake(const ShaderOutput<void*>& shader, const std::vector<std::string>* resource_map, float* output) = 0;
    virtual void runPass(const ShaderOutput<void*>& shader) = 0;
};

// src/runtime/device/ICompilerDevice.h:6-16
class ICompilerDevice {
public:
    virtual ~ICompilerDevice() = default;
    struct Settings { int OptimizationLevel = 3; bool Verbose = false; };
    virtual bool compile(const Settings& settings, const std::string& script) const = 0;
    virtual void* compileAndGet(const Settings& settings, const std::string& script, const std::string& function) const = 0;
};

// src/runtime/device/IDeviceInterface.h:9-17
class IDeviceInterface {
public:
    virtual ~IDeviceInterface() = default;
    virtual Build::Version getVersion() const = 0;
    virtual GPUArchitecture getArchitecture() const = 0;   // the reference returns std::variant<CPU, GPU>; a GPU plugin always holds the GPU alternative
    virtual IRenderDevice* createRenderDevice(const IRenderDevice::SetupSettings& settings) const = 0;
    virtual ICompilerDevice* createCompilerDevice() const = 0;
};

}  // namespace IG
#endif
