/* TEST INFRASTRUCTURE -- part of the CPU oracle, never linked into the product.
 *
 * Deterministic single-precision transcendentals used by the oracle.
 *
 * Why they exist: the reference calls the platform math library (`math_builtins::sin/cos/acos/atan2`,
 * e.g. src/artic/core/sampling.art:13-21, src/artic/light/area.art:152-159,175-178,
 * src/artic/core/warp.art:63-91, src/artic/shapes/sphere.art:1-6) and is built with -ffast-math
 * (CMakeLists.txt:93-98), so its last-ulp behaviour is not pinned by anything in the reference tree
 * (SURVEY.md 8c). The radiance-parity bar (1e-4 relative L2) needs every discrete decision of a path
 * (Russian roulette, Fresnel choice, light pick) to agree between CPU and GPU, which a 1-ulp difference
 * in sin() can flip. Both sides therefore evaluate the SAME published polynomial kernels, built only
 * from IEEE-754 +,-,*,/,sqrt and fused multiply-add, all of which are correctly rounded on x86-64 and
 * on sm_100a. The CUDA side has its own copy (ignis_b200/csrc/detmath.cuh); tests/test_detmath.py
 * checks this file against libm (<= 4 ulp) and the GPU copy against this file (bit-exact).
 *
 * Algorithms (published, restated):
 *  - sin/cos: Cody-Waite 3-constant reduction by pi/2 + minimax kernels on [-pi/4, pi/4]
 *    (coefficients as in N. Juffa's public-domain sincosf, also the shape of CUDA's own sinf).
 *  - asin/acos/atan: S. Moshier, Cephes Mathematical Library 2.8, single precision (asinf.c, atanf.c).
 */
#ifndef IGO_DETMATH_H
#define IGO_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float dm_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }

static inline int32_t dm_f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

/* sin and cos of x, |x| < ~1e4 */
static inline void dm_sincosf(float x, float* sp, float* cp)
{
    /* j = nearest integer to x * 2/pi, via the 1.5*2^23 magic constant */
    float j = dm_fma(x, 0.636619747f, 12582912.0f);
    const int32_t q = dm_f2i(j);
    j = j - 12582912.0f;
    float r = dm_fma(j, -1.57079601e+00f, x);
    r = dm_fma(j, -3.13916473e-07f, r);
    r = dm_fma(j, -5.39030253e-15f, r);
    const float s = r * r;
    /* cosine kernel */
    float c = 2.44677067e-5f;
    c = dm_fma(c, s, -1.38877297e-3f);
    c = dm_fma(c, s, 4.16666567e-2f);
    c = dm_fma(c, s, -5.00000000e-1f);
    c = dm_fma(c, s, 1.00000000e+0f);
    /* sine kernel */
    float p = 2.86567956e-6f;
    p = dm_fma(p, s, -1.98559923e-4f);
    p = dm_fma(p, s, 8.33338592e-3f);
    p = dm_fma(p, s, -1.66666672e-1f);
    const float t = r * s;
    p = dm_fma(p, t, r);
    float sn = p, cs = c;
    if (q & 1) { sn = c; cs = p; }
    if (q & 2) sn = -sn;
    if ((q + 1) & 2) cs = -cs;
    *sp = sn;
    *cp = cs;
}

static inline float dm_sinf(float x) { float s, c; dm_sincosf(x, &s, &c); return s; }
static inline float dm_cosf(float x) { float s, c; dm_sincosf(x, &s, &c); return c; }

/* Cephes asinf kernel: asin(x) for 0 <= x <= 0.5 */
static inline float dm_asin_kernel(float x)
{
    const float z = x * x;
    float p = 4.2163199048e-2f;
    p = dm_fma(p, z, 2.4181311049e-2f);
    p = dm_fma(p, z, 4.5470025998e-2f);
    p = dm_fma(p, z, 7.4953002686e-2f);
    p = dm_fma(p, z, 1.6666752422e-1f);
    return dm_fma(p * z, x, x);
}

/* acos(x), x in [-1, 1] (Cephes acosf) */
static inline float dm_acosf(float x)
{
    if (x < -0.5f) {
        const float w = sqrtf(0.5f * (1.0f + x));
        return 3.14159265358979323846f - 2.0f * dm_asin_kernel(w);
    }
    if (x > 0.5f) {
        const float w = sqrtf(0.5f * (1.0f - x));
        return 2.0f * dm_asin_kernel(w);
    }
    const float a = fabsf(x);
    const float r = dm_asin_kernel(a);
    return 1.5707963267948966192f - (x < 0.0f ? -r : r);
}

/* atan(x) (Cephes atanf) */
static inline float dm_atanf(float xx)
{
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) {        /* tan 3pi/8 */
        y = 1.5707963267948966192f;
        x = -(1.0f / x);
    } else if (x > 0.4142135623730950f) { /* tan pi/8 */
        y = 0.7853981633974483096f;
        x = (x - 1.0f) / (x + 1.0f);
    } else {
        y = 0.0f;
    }
    const float z = x * x;
    float p = 8.05374449538e-2f;
    p = dm_fma(p, z, -1.38776856032e-1f);
    p = dm_fma(p, z, 1.99777106478e-1f);
    p = dm_fma(p, z, -3.33329491539e-1f);
    y = y + dm_fma(p * z, x, x);
    return xx < 0.0f ? -y : y;
}

/* atan2(y, x) (Cephes atan2f quadrant logic) */
static inline float dm_atan2f(float y, float x)
{
    const float pi = 3.14159265358979323846f;
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966192f;
        if (y < 0.0f) return -1.5707963267948966192f;
        return 0.0f;
    }
    if (y == 0.0f) return x < 0.0f ? pi : 0.0f;
    const float z = dm_atanf(y / x);
    if (x < 0.0f) return y < 0.0f ? z - pi : z + pi;
    return z;
}

#endif
