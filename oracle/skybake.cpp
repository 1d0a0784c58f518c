// TEST / FIXTURE INFRASTRUCTURE -- bakes the image the reference's `sky` light renders from (src/runtime/light/SkyLight.cpp:28-47).
// Compiled TOGETHER WITH the reference's own sources where they lie under /root/reference (oracle/Makefile: _ref/skybake):
//   src/runtime/skysun/model/ArHosekSkyModel.cpp   the Hosek-Wilkie model and its coefficient tables
//   src/runtime/skysun/SunLocation.cpp             sun position from date, time and place
// Only the 25-line sampling loop of src/runtime/skysun/SkyModel.cpp:9-52 is restated here (that file needs the reference's image and
// logging classes). Output: RES_AZ x RES_EL x 3 float32, row 0 = zenith, on stdout. The sky model is loader-side data production, not on
// the hot path: the device receives the image (include/igb200.h IGB200_LIGHT_ENV_TEXTURED).
//   skybake ground_r ground_g ground_b turbidity (ea <elevation> <azimuth> | dir <x> <y> <z> | time <Y> <M> <D> <h> <m> <s> <lat> <lon> <tz>)
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "skysun/SunLocation.h"
#include "skysun/model/ArHosekSkyModel.h"

using namespace IG;

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const float ground[3] = {(float)atof(argv[1]), (float)atof(argv[2]), (float)atof(argv[3])};
    const float turbidity = (float)atof(argv[4]);
    ElevationAzimuth ea{0, 0};
    if (!strcmp(argv[5], "ea") && argc >= 8) ea = ElevationAzimuth{(float)atof(argv[6]), (float)atof(argv[7])};
    else if (!strcmp(argv[5], "dir") && argc >= 9) {   // LoaderUtils::getEA: direction normalised, then fromDirectionYUp
        float d[3] = {(float)atof(argv[6]), (float)atof(argv[7]), (float)atof(argv[8])};
        const float l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        ea = ElevationAzimuth::fromDirectionYUp(Vector3f(d[0] / l, d[1] / l, d[2] / l));
    } else if (!strcmp(argv[5], "time") && argc >= 15) {
        TimePoint tp(atoi(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]), atoi(argv[10]), (float)atof(argv[11]));
        MapLocation loc((float)atof(argv[13]), (float)atof(argv[12]), (float)atof(argv[14]));
        ea = computeSunEA(tp, loc);
    } else return 2;
    const size_t az_n = 512, el_n = 256;   // skysun/SkySunConfig.h:8-9
    // ---- skysun/SkyModel.cpp:13-52
    const float solar_elevation = Pi2 - ea.Elevation;
    const float sun_se = std::sin(solar_elevation), sun_ce = std::cos(solar_elevation);
    std::array<ArHosekSkyModelState*, 3> states;
    for (size_t k = 0; k < 3; ++k) states[k] = arhosek_rgb_skymodelstate_alloc_init(turbidity, ground[k], solar_elevation);
    std::vector<float> data(el_n * az_n * 3, 0.0f);
    for (size_t y = 0; y < el_n; ++y) {
        const float theta = ELEVATION_RANGE * y / (float)el_n;
        const float st = std::sin(theta), ct = std::cos(theta);
        for (size_t x = 0; x < az_n; ++x) {
            float azimuth = AZIMUTH_RANGE * x / (float)az_n - Pi4;
            if (azimuth < 0) azimuth += 2 * Pi;
            const float cosGamma = ct * sun_ce + st * sun_se * std::cos(azimuth - ea.Azimuth);
            const float gamma = std::acos(std::min(1.0f, std::max(-1.0f, cosGamma)));
            for (size_t k = 0; k < 3; ++k) {
                constexpr float CIEYSum = 106.856980f;
                const float radiance = (float)arhosek_tristim_skymodel_radiance(states[k], theta, gamma, (int)k) / CIEYSum;
                data[(y * az_n + x) * 3 + k] = std::max(0.0f, radiance);
            }
        }
    }
    for (auto* s : states) arhosekskymodelstate_free(s);
    fprintf(stderr, "elevation %.9g azimuth %.9g\n", ea.Elevation, ea.Azimuth);
    fwrite(data.data(), sizeof(float), data.size(), stdout);
    return 0;
}
