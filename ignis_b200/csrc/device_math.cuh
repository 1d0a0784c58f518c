// Device-side arithmetic of the hot path: vectors, matrices, RNG, deterministic transcendentals.
//
// Compiled with -fmad=false: the ONLY fused multiply-adds are the explicit __fmaf_rn calls below, placed where the
// reference itself writes fmaf (vec*_dot: src/artic/core/vector.art:98-100; sum_of_prod: core/common.art:257-272;
// spherical-rectangle sampling: light/area.art:175-190) plus the ray/box slabs (DESIGN.md "Numerics").
// Every other expression keeps the reference's operation order so that results are reproducible bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace igb {

// src/artic/core/common.art:3-8
#define IGB_FLT_EPS 1.1920928955e-07f
#define IGB_FLT_MAX 3.4028234664e+38f
#define IGB_FLT_PI 3.14159265359f
#define IGB_FLT_INV_PI 0.31830988618379067154f

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 mulf(V3 a, float t) { return v3(a.x * t, a.y * t, a.z * t); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fma_(a.x, b.x, fma_(a.y, b.y, a.z * b.z)); }   // vector.art:99
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float len2(V3 a) { return dot(a, a); }
__device__ __forceinline__ float len(V3 a) { return sqrtf(len2(a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return mulf(a, 1.0f / len(a)); }
__device__ __forceinline__ float lerp2(float a, float b, float c, float k1, float k2) { return (1 - k1 - k2) * a + k1 * b + k2 * c; }  // common.art:255
__device__ __forceinline__ V3 lerp2(V3 a, V3 b, V3 c, float u, float v) { return v3(lerp2(a.x, b.x, c.x, u, v), lerp2(a.y, b.y, c.y, u, v), lerp2(a.z, b.z, c.z, u, v)); }
__device__ __forceinline__ float lerp1(float a, float b, float k) { return (1 - k) * a + k * b; }

__device__ __forceinline__ float prodsign(float x, float y) { return __uint_as_float(__float_as_uint(x) ^ (__float_as_uint(y) & 0x80000000u)); }
__device__ __forceinline__ float safe_rcp(float x) { return ((x > 0 ? x : -x) < 1e-8f) ? prodsign(IGB_FLT_MAX, x) : 1.0f / x; }   // common.art:210-213
__device__ __forceinline__ float clampf(float v, float l, float u) { return fminf(u, fmaxf(l, v)); }
__device__ __forceinline__ float safe_div(float a, float b) { return fabsf(b) <= IGB_FLT_EPS ? 0.0f : a / b; }
__device__ __forceinline__ float safe_sqrt(float a) { return sqrtf(fmaxf(0.0f, a)); }
__device__ __forceinline__ float sum_of_prod(float a, float b, float c, float d) { float cd = c * d; float s = fma_(a, b, cd); float e = fma_(c, d, -cd); return s + e; }
__device__ __forceinline__ float positive_cos(V3 a, V3 b) { float c = dot(a, b); return c >= 0 ? c : 0.0f; }

struct M33 { V3 c0, c1, c2; };           // column major
struct M34 { V3 c0, c1, c2, c3; };
__device__ __forceinline__ V3 m33_mul(const M33& m, V3 v) {  // matrix.art:105-108
    return v3(dot(v3(m.c0.x, m.c1.x, m.c2.x), v), dot(v3(m.c0.y, m.c1.y, m.c2.y), v), dot(v3(m.c0.z, m.c1.z, m.c2.z), v));
}
__device__ __forceinline__ float dot4(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw) {
    return fma_(ax, bx, fma_(ay, by, fma_(az, bz, aw * bw)));
}
// rows given as float4 (m_r0, m_r1, m_r2, m_r3): matrix.art:115-118,246-247
__device__ __forceinline__ V3 xform_point(float4 r0, float4 r1, float4 r2, V3 v) {
    return v3(dot4(r0.x, r0.y, r0.z, r0.w, v.x, v.y, v.z, 1), dot4(r1.x, r1.y, r1.z, r1.w, v.x, v.y, v.z, 1), dot4(r2.x, r2.y, r2.z, r2.w, v.x, v.y, v.z, 1));
}
__device__ __forceinline__ V3 xform_dir(float4 r0, float4 r1, float4 r2, V3 v) {
    return v3(dot4(r0.x, r0.y, r0.z, r0.w, v.x, v.y, v.z, 0), dot4(r1.x, r1.y, r1.z, r1.w, v.x, v.y, v.z, 0), dot4(r2.x, r2.y, r2.z, r2.w, v.x, v.y, v.z, 0));
}
__device__ __forceinline__ M33 make_orthonormal(V3 n) {  // matrix.art:24-32
    const float sign = copysignf(1.0f, n.z);
    const float a = -1 / (sign + n.z);
    const float b = n.x * n.y * a;
    M33 m;
    m.c0 = v3(1 + sign * n.x * n.x * a, sign * b, -sign * n.x);
    m.c1 = v3(b, sign + n.y * n.y * a, -n.y);
    m.c2 = n;
    return m;
}

// ---- RNG: src/artic/core/random.art:7-24,34-43,45-87
__device__ __forceinline__ uint32_t hash_combine(uint32_t h, uint32_t d) {
    h = (h * 16777619u) ^ (d & 0xFF);
    h = (h * 16777619u) ^ ((d >> 8) & 0xFF);
    h = (h * 16777619u) ^ ((d >> 16) & 0xFF);
    h = (h * 16777619u) ^ ((d >> 24) & 0xFF);
    return h;
}
__device__ __forceinline__ uint32_t tea4(uint32_t v0, uint32_t v1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v1;
}
__device__ __forceinline__ uint32_t random_seed(int sample, int iter, int frame, int x, int y, int user) {
    uint32_t h = 0x811C9DC5u;
    h = hash_combine(h, (uint32_t)sample);
    h = hash_combine(h, (uint32_t)iter);
    h = hash_combine(h, (uint32_t)frame);
    h = hash_combine(h, (uint32_t)x);
    h = hash_combine(h, (uint32_t)y);
    h = hash_combine(h, (uint32_t)user);
    return h;
}
struct Rng {
    uint32_t seed, counter;
    __device__ __forceinline__ uint32_t next_u32() { return tea4(seed, counter++); }
    __device__ __forceinline__ float next_f32() { const uint32_t x = next_u32(); return __uint_as_float((x & 0x7FFFFFu) | 0x3F800000u) - 1; }
    __device__ __forceinline__ int next_i32(int s, int e) {
        const uint32_t range = (uint32_t)(e - s);
        if (range == 0xFFFFFFFFu) return (int)next_u32() + s;
        const uint32_t erange = range + 1, scaling = 0xFFFFFFFFu / erange, past = erange * scaling;
        uint32_t ret = next_u32();
        while (ret >= past) ret = next_u32();
        return (int)(ret / scaling) + s;
    }
};

// ---- deterministic transcendentals: the same kernels as oracle/detmath.h (see its header for provenance)
__device__ __forceinline__ void dm_sincosf(float x, float* sp, float* cp) {
    float j = fma_(x, 0.636619747f, 12582912.0f);
    const int q = __float_as_int(j);
    j = j - 12582912.0f;
    float r = fma_(j, -1.57079601e+00f, x);
    r = fma_(j, -3.13916473e-07f, r);
    r = fma_(j, -5.39030253e-15f, r);
    const float s = r * r;
    float c = 2.44677067e-5f;
    c = fma_(c, s, -1.38877297e-3f);
    c = fma_(c, s, 4.16666567e-2f);
    c = fma_(c, s, -5.00000000e-1f);
    c = fma_(c, s, 1.00000000e+0f);
    float p = 2.86567956e-6f;
    p = fma_(p, s, -1.98559923e-4f);
    p = fma_(p, s, 8.33338592e-3f);
    p = fma_(p, s, -1.66666672e-1f);
    const float t = r * s;
    p = fma_(p, t, r);
    float sn = p, cs = c;
    if (q & 1) { sn = c; cs = p; }
    if (q & 2) sn = -sn;
    if ((q + 1) & 2) cs = -cs;
    *sp = sn;
    *cp = cs;
}
__device__ __forceinline__ float dm_asin_kernel(float x) {
    const float z = x * x;
    float p = 4.2163199048e-2f;
    p = fma_(p, z, 2.4181311049e-2f);
    p = fma_(p, z, 4.5470025998e-2f);
    p = fma_(p, z, 7.4953002686e-2f);
    p = fma_(p, z, 1.6666752422e-1f);
    return fma_(p * z, x, x);
}
__device__ __forceinline__ float dm_acosf(float x) {
    if (x < -0.5f) { const float w = sqrtf(0.5f * (1.0f + x)); return 3.14159265358979323846f - 2.0f * dm_asin_kernel(w); }
    if (x > 0.5f) { const float w = sqrtf(0.5f * (1.0f - x)); return 2.0f * dm_asin_kernel(w); }
    const float a = fabsf(x);
    const float r = dm_asin_kernel(a);
    return 1.5707963267948966192f - (x < 0.0f ? -r : r);
}
__device__ __forceinline__ float dm_atanf(float xx) {
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966192f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483096f; x = (x - 1.0f) / (x + 1.0f); }
    else { y = 0.0f; }
    const float z = x * x;
    float p = 8.05374449538e-2f;
    p = fma_(p, z, -1.38776856032e-1f);
    p = fma_(p, z, 1.99777106478e-1f);
    p = fma_(p, z, -3.33329491539e-1f);
    y = y + fma_(p * z, x, x);
    return xx < 0.0f ? -y : y;
}
__device__ __forceinline__ float dm_atan2f(float y, float x) {
    const float pi = 3.14159265358979323846f;
    if (x == 0.0f) { if (y > 0.0f) return 1.5707963267948966192f; if (y < 0.0f) return -1.5707963267948966192f; return 0.0f; }
    if (y == 0.0f) return x < 0.0f ? pi : 0.0f;
    const float z = dm_atanf(y / x);
    if (x < 0.0f) return y < 0.0f ? z - pi : z + pi;
    return z;
}

}  // namespace igb
