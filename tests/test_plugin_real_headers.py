"""The C++ plugin layer (ignis_b200/csrc/host) type-checked against the REFERENCE'S OWN interface headers.

`g++ -std=c++20 -fsyntax-only -DIGB200_WITH_IGNIS -I<reference>/src/runtime`: ig_mirror.h then includes the real
device/IDeviceInterface.h (-> IRenderDevice.h, ICompilerDevice.h, Target.h, TechniqueVariant.h, ParameterSet.h,
RuntimeStructs.h, config/Build.h), table/SceneDatabase.h and Statistics.h instead of its restatement, so every `override`
in b200_device.h is checked against the real pure-virtual it implements (src/runtime/device/IDeviceInterface.h:9-17,
IRenderDevice.h:14-81, ICompilerDevice.h:6-16; the reference's own plugin: src/device/Interface.cpp:16-76). The only
stand-in is Eigen (absent from this container): tests/stubs/Eigen is a syntax-only subset. Skipped where the reference tree is
absent (the GPU box) -- the normal build of the same sources against ig_mirror.h runs everywhere.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/runtime"
HOST = os.path.join(ROOT, "ignis_b200", "csrc", "host")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("g++") is None, reason="needs the reference's headers and g++")


def syntax_check(source, extra=()):
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Wextra", "-Werror=overloaded-virtual", "-Werror=suggest-override", "-DIGB200_WITH_IGNIS",
           "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + REF, *extra, source]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.mark.parametrize("source", ["b200_device.cpp", "script_recognizer.cpp", "image_io.cpp", "host_capi.cpp"])
def test_host_layer_compiles_against_the_reference_headers(source):
    r = syntax_check(os.path.join(HOST, source))
    assert r.returncode == 0, r.stderr[-4000:]


def test_plugin_classes_are_concrete_implementations_of_the_real_interfaces(tmp_path):
    # instantiable (no pure virtual left over) and convertible to the reference's base classes; the version the plugin reports is the
    # one DeviceManager.cpp:180-194 compares with the runtime's
    src = tmp_path / "probe.cpp"
    src.write_text(
        '#include "b200_device.h"\n'
        '#include <type_traits>\n'
        'static_assert(std::is_base_of_v<IG::IRenderDevice, igbh::B200Device> && !std::is_abstract_v<igbh::B200Device>);\n'
        'static_assert(std::is_base_of_v<IG::ICompilerDevice, igbh::B200CompilerDevice> && !std::is_abstract_v<igbh::B200CompilerDevice>);\n'
        'static_assert(std::is_base_of_v<IG::IDeviceInterface, igbh::B200DeviceInterface> && !std::is_abstract_v<igbh::B200DeviceInterface>);\n'
        'static_assert(std::is_same_v<decltype(ig_get_interface()), const IG::IDeviceInterface*>);\n'
        'static_assert(std::is_same_v<decltype(std::declval<igbh::B200DeviceInterface>().getArchitecture()), IG::TargetArchitecture>);\n'
        'static_assert(std::is_same_v<IG::Vector3f, Eigen::Vector3f>, "the real IG_Config.h is in use, not the mirror");\n'
        'static_assert(IGB200_VERSION_MAJOR == 0 && IGB200_VERSION_MINOR == 3);\n')
    r = syntax_check(str(src), extra=["-I" + HOST])
    assert r.returncode == 0, r.stderr[-4000:]


def test_a_signature_mismatch_is_caught(tmp_path):
    # the check has teeth: the round-1 declaration (`GPUArchitecture getArchitecture()`) must not compile against the real header
    text = open(os.path.join(HOST, "b200_device.h")).read()
    bad = text.replace("IG::TargetArchitecture getArchitecture() const override { return IG::TargetArchitecture{IG::GPUArchitecture::Nvidia}; }",
                       "IG::GPUArchitecture getArchitecture() const override { return IG::GPUArchitecture::Nvidia; }")
    assert bad != text
    d = tmp_path / "ignis_b200" / "csrc" / "host"   # same depth as in the repo, so that the relative include of include/igb200.h resolves
    shutil.copytree(HOST, d)
    os.symlink(os.path.join(ROOT, "include"), tmp_path / "include")
    (d / "b200_device.h").write_text(bad)
    r = syntax_check(str(d / "b200_device.cpp"))
    assert r.returncode != 0 and "getArchitecture" in r.stderr
