// ig_mirror.h -- the part of the reference's device-plugin contract that the B200 host layer implements, restated.
//
// The reference's plugin boundary is a C++ vtable: `ig_get_interface()` returns an `IDeviceInterface`
// (src/runtime/device/IDeviceInterface.h:9-17) that creates an `IRenderDevice` (src/runtime/device/IRenderDevice.h:14-81)
// and an `ICompilerDevice` (src/runtime/device/ICompilerDevice.h:6-16). Its real headers drag in Eigen, TBB and the
// generated `generated_interface.h`, none of which exist in this container, so this file restates ONLY the types that
// cross the boundary on the `path` hot path, with the reference's names, member names and meaning. Inside the reference's
// build tree, define IGB200_WITH_IGNIS and the real headers are used instead; `namespace IG` is the
// same in both cases and the restated types offer the members of the real ones that the host layer touches (Eigen's
// `v(i)` / `x()` accessors, `Statistics::reset / increase`, `TargetArchitecture`), so b200_device.cpp and
// script_recognizer.cpp compile unchanged against either -- tests/test_plugin_real_headers.py type-checks them against the
// reference's own headers (g++ -std=c++20 -fsyntax-only -DIGB200_WITH_IGNIS -I<reference>/src/runtime).
#pragma once

#ifdef IGB200_WITH_IGNIS
#include "device/IDeviceInterface.h"
#include "table/SceneDatabase.h"
#include "Statistics.h"
#else

#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <variant>
#include <vector>

namespace IG {

using uint8 = uint8_t;
using int32 = int32_t;
using uint32 = uint32_t;
using uint64 = uint64_t;

// Stand-ins for the Eigen vectors of the reference (same storage: packed floats; the accessors are Eigen's: v(i), v[i], x() ...)
template <int N> struct VecNf {
    float m[N] = {};
    VecNf() = default;
    VecNf(float a, float b) : m{a, b} { static_assert(N == 2, "size"); }
    VecNf(float a, float b, float c) : m{a, b, c} { static_assert(N == 3, "size"); }
    VecNf(float a, float b, float c, float d) : m{a, b, c, d} { static_assert(N == 4, "size"); }
    static VecNf Zero() { return VecNf(); }
    float& operator()(std::ptrdiff_t i) { return m[i]; }
    float operator()(std::ptrdiff_t i) const { return m[i]; }
    float& operator[](std::ptrdiff_t i) { return m[i]; }
    float operator[](std::ptrdiff_t i) const { return m[i]; }
    float x() const { return m[0]; } float y() const { return m[1]; } float z() const { return m[2]; } float w() const { return m[3]; }
};
using Vector2f = VecNf<2>; using Vector3f = VecNf<3>; using Vector4f = VecNf<4>;

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
enum class CPUArchitecture { ARM, X86, Unknown };
using TargetArchitecture = std::variant<CPUArchitecture, GPUArchitecture>;
struct Target {
    bool gpu = true; GPUArchitecture arch = GPUArchitecture::Nvidia; size_t dev = 0;
    bool isGPU() const { return gpu; }
    GPUArchitecture gpuArchitecture() const { return arch; }
    size_t device() const { return dev; }
    void setDevice(size_t d) { dev = d; }
    static Target makeGPU(GPUArchitecture a, size_t device) { Target t; t.gpu = true; t.arch = a; t.dev = device; return t; }
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

// src/runtime/Statistics.h:57-64,69-105 (the ray counters; the rest of the class is host-side bookkeeping). As in the reference the
// counters are private: the runtime only ever merges (`add`) and prints (`dump`) them.
enum class Quantity { CameraRayCount = 0, ShadowRayCount, BounceRayCount, _COUNT };
class Statistics {
public:
    void reset() { *this = Statistics(); }
    void increase(Quantity quantity, uint64 value) { mQuantities[(size_t)quantity] += value; }
    void add(const Statistics& other) { for (size_t i = 0; i < mQuantities.size(); ++i) mQuantities[i] += other.mQuantities[i]; }
private:
    std::array<uint64, (size_t)Quantity::_COUNT> mQuantities{};
};

struct TonemapSettings; struct ImageInfoSettings;
struct ImageInfoOutput { float Min, Max, Average, SoftMin, SoftMax, Median; int InfCount, NaNCount, NegCount; };   // RuntimeStructs.h:27-37

namespace Build {   // src/runtime/config/Build.h:6-13
struct Version { uint32 Major, Minor; uint32 asNumber() const { return (Major << 8) | Minor; } };
inline bool operator==(const Version& a, const Version& b) { return a.asNumber() == b.asNumber(); }
}

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
    virtual TargetArchitecture getArchitecture() const = 0;
    virtual IRenderDevice* createRenderDevice(const IRenderDevice::SetupSettings& settings) const = 0;
    virtual ICompilerDevice* createCompilerDevice() const = 0;
};

}  // namespace IG
#endif
