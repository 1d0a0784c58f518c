// TEST / FIXTURE INFRASTRUCTURE. Stand-in for the reference's src/runtime/IG_Config.h (which pulls in Eigen and TBB, absent here) that is
// just enough to compile src/runtime/skysun/SunLocation.cpp and ElevationAzimuth.h where they lie under /root/reference
// (oracle/Makefile target _ref/skybake). Constants as src/runtime/IG_Config.h:255-268.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <ctime>
#define IG_LIB
namespace IG {
constexpr float Pi = 3.14159265358979323846f;
constexpr float Pi2 = 1.57079632679489661923f;
constexpr float Pi4 = 0.78539816339744830961f;
constexpr float Deg2Rad = Pi / 180.0f;
struct Vector3f {
    float v[3];
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    float operator()(int i) const { return v[i]; }
};
}  // namespace IG
