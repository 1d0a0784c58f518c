// b200_device.h -- the reference's device-plugin classes implemented over the C ABI (include/igb200.h).
//
//   B200Device          : IG::IRenderDevice    replaces src/device/Device.cpp (`class Interface` :143-1605, `Device` :1607-1860)
//   B200CompilerDevice  : IG::ICompilerDevice  replaces src/device/Compiler.cpp (anydsl_compile) with script_recognizer
//   B200DeviceInterface : IG::IDeviceInterface replaces src/device/Interface.cpp:16-68
//   ig_get_interface()                          replaces src/device/Interface.cpp:70-76
//
// Method for method the behaviour is the reference's; what differs is documented at each member in b200_device.cpp.
#pragma once

#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/igb200.h"
#include "ig_mirror.h"
#include "script_recognizer.h"

namespace igbh {

class B200CompilerDevice : public IG::ICompilerDevice {
public:
    bool compile(const Settings& settings, const std::string& script) const override;
    void* compileAndGet(const Settings& settings, const std::string& script, const std::string& function) const override;
private:
    // handles stay valid for the lifetime of the compiler device (the runtime keeps the void* in its shader sets)
    mutable std::mutex mMutex;
    mutable std::vector<std::unique_ptr<StageDescriptor>> mStages;
};

class B200Device : public IG::IRenderDevice {
public:
    explicit B200Device(const SetupSettings& settings);
    ~B200Device() override;
    bool valid() const { return mCtx != nullptr; }
    int gpuCount() const { return (int)mCtxs.size(); }

    void assignScene(const SceneSettings& settings) override;
    void render(const IG::TechniqueVariantShaderSet& shader_set, const RenderSettings& settings, IG::ParameterSet* parameter_set) override;
    void resize(size_t width, size_t height) override;
    void releaseAll() override;
    IG::Target target() const override { return mSetup.target; }
    size_t framebufferWidth() const override { return mWidth; }
    size_t framebufferHeight() const override { return mHeight; }
    bool isInteractive() const override { return mSetup.IsInteractive; }
    AOVAccessor getFramebufferForHost(const std::string& name, bool sync = true) override;
    AOVAccessor getFramebufferForDevice(const std::string& name, bool sync = true) override;
    void clearFramebuffer(const std::string& name) override;
    void clearAllFramebuffer() override;
    void syncFramebufferHostToDevice(const std::string& name) override;
    void syncAllFramebufferHostToDevice() override;
    size_t getBufferSizeInBytes(const std::string& name) override;
    bool copyBufferToHost(const std::string& name, void* buffer, size_t maxSizeByte) override;
    BufferAccessor getBufferForDevice(const std::string& name) override;
    const IG::Statistics* getStatistics() override;
    void tonemap(uint32_t*, const IG::TonemapSettings&) override;
    IG::ImageInfoOutput imageinfo(const IG::ImageInfoSettings&) override;
    void bake(const IG::ShaderOutput<void*>& shader, const std::vector<std::string>* resource_map, float* output) override;
    void runPass(const IG::ShaderOutput<void*>& shader) override;

    // device-specific (no reference counterpart): multi-GPU tile partition, see include/igb200.h
    bool setPartition(int rank, int world, int tile);
    bool rayCounters(uint64_t out[3], double* render_ms = nullptr);   // camera, shadow, bounce rays since the last reset
    igb200_ctx* context() { return mCtx; }
    const std::string& lastError() const { return mError; }

private:
    bool uploadScene(const IG::TechniqueVariantShaderSet& shader_set, const IG::ParameterSet* global);
    void error(const std::string& what);
    // Runs f(context, rank) on every GPU of the device -- on one thread per GPU when there are several (NCCL calls of the ranks of one
    // process must not be serialised: a rank's send completes only once the root's receive is posted, and connections are made lazily by
    // both sides). Returns false and reports the first failure otherwise.
    bool forAll(const std::function<int(igb200_ctx*, int)>& f, const char* what);
    void destroyAll();

    SetupSettings mSetup;
    SceneSettings mScene{};
    igb200_ctx* mCtx = nullptr;              // rank 0: the GPU the frame is assembled on
    std::vector<igb200_ctx*> mCtxs;          // one context per GPU (IGB200_GPUS), mCtxs[0] == mCtx
    size_t mWidth = 0, mHeight = 0;
    bool mSceneDirty = true;
    int mStdAovs = -1;                       // Normals / Albedo AOVs currently enabled on the device (-1: not set yet)
    std::vector<uint8_t> mDescriptorBytes;   // materials + lights + camera + technique of the scene on the device
    ImageCache mFiles;                       // decoded image / buffer files the stage text names (dropped with the scene)
    std::vector<float> mHostFramebuffer;     // what getFramebufferForHost handed out, for syncFramebufferHostToDevice
    std::unordered_map<std::string, float*> mHostPtrs;   // per AOV ("" = Color): the context-owned host buffer last handed out
    IG::Statistics mStats;
    std::string mError;
};

class B200DeviceInterface : public IG::IDeviceInterface {
public:
    IG::Build::Version getVersion() const override { return IG::Build::Version{IGB200_VERSION_MAJOR, IGB200_VERSION_MINOR}; }
    IG::TargetArchitecture getArchitecture() const override { return IG::TargetArchitecture{IG::GPUArchitecture::Nvidia}; }   // it replaces ig_device_cuda (Target.h:7-21)
    IG::IRenderDevice* createRenderDevice(const IG::IRenderDevice::SetupSettings& settings) const override;
    IG::ICompilerDevice* createCompilerDevice() const override { return new B200CompilerDevice(); }
};

}  // namespace igbh

extern "C" const IG::IDeviceInterface* ig_get_interface();
