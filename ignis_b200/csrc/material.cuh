// Textures, microfacets, normal / bump mapping, the general BSDF and the textured environment lights of the shade phase
// (the "full" shade kernels only: shade.cuh shade_record<true>). Every function follows the reference expression by expression
// (paths relative to the reference's src/artic) so that results equal the oracle's bit for bit; compiled with -fmad=false.
#pragma once

#include "types.cuh"

namespace igb {

// ---- textures: texture/checkerboard.art:4-13, texture/image.art:9-153, driver/image.art:9-34 ------------------------------------
__device__ __forceinline__ float fract_(float x) { return x - floorf(x); }                                                   // core/math.art:74
__device__ __forceinline__ float wrap_(float v, float mn, float mx) { const float range = mx - mn; return range <= IGB_FLT_EPS ? mn : v - (range * floorf((v - mn) / range)); }   // core/math.art:88-91
__device__ __forceinline__ float4 f4lerp(float4 a, float4 b, float t) { return make_float4((1 - t) * a.x + t * b.x, (1 - t) * a.y + t * b.y, (1 - t) * a.z + t * b.z, (1 - t) * a.w + t * b.w); }   // core/color.art:17-21
__device__ __forceinline__ float4 f4mulf(float4 a, float f) { return make_float4(a.x * f, a.y * f, a.z * f, a.w * f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ int border_index(int mode, int x, int w) {
    if (mode == 1) return x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
    if (mode == 2) { const int t = x < 0 ? -1 - x : x; const int i = t / w; const int k = t - i * w; return (i & 1) == 0 ? w - 1 - k : k; }
    const int t = x % w; return t < 0 ? t + w : t;
}
// im = (format, width, height, first 32-bit word of the pixels inside image_data)
__device__ __forceinline__ float4 image_pixel(const DevScene& sc, int4 im, int x, int y) {
    const size_t i = (size_t)y * (size_t)im.y + (size_t)x;
    if (im.x == 0) { const uint32_t p = __ldg(sc.image_data + im.w + i); return make_float4((float)(p & 0xFFu) / 255, (float)((p >> 8) & 0xFFu) / 255, (float)((p >> 16) & 0xFFu) / 255, (float)(p >> 24) / 255); }
    if (im.x == 1) { const float g = (float)__ldg(reinterpret_cast<const unsigned char*>(sc.image_data + im.w) + i) / 255; return make_float4(g, g, g, 1); }
    return __ldg(reinterpret_cast<const float4*>(sc.image_data + im.w) + i);
}
__device__ __forceinline__ float cubic_w0(float a) { return (a * (a * (-a + 3) - 3) + 1) / 6; }
__device__ __forceinline__ float cubic_w1(float a) { return (a * a * (3 * a - 6) + 4) / 6; }
__device__ __forceinline__ float cubic_w2(float a) { return (a * (a * (-3 * a + 3) + 3) + 1) / 6; }
__device__ __forceinline__ float cubic_w3(float a) { return (a * a * a) / 6; }
__device__ __forceinline__ float cubic_g0(float a) { return cubic_w0(a) + cubic_w1(a); }
__device__ __forceinline__ float cubic_g1(float a) { return cubic_w2(a) + cubic_w3(a); }
__device__ __forceinline__ float cubic_h0(float a) { return (cubic_w1(a) / cubic_g0(a)) - 1; }
__device__ __forceinline__ float cubic_h1(float a) { return (cubic_w3(a) / cubic_g1(a)) + 1; }

__device__ __noinline__ float4 image_filter(const DevScene& sc, int image, int filter, int bu, int bv, float uvx, float uvy) {
    const int4 im = __ldg(sc.images + image);
    if (filter == 0) {
        const float u = uvx * (float)im.y, v = uvy * (float)im.z;
        return image_pixel(sc, im, border_index(bu, (int)floorf(u), im.y), border_index(bv, (int)floorf(v), im.z));
    }
    const float u = uvx * (float)im.y - 0.5f, v = uvy * (float)im.z - 0.5f;
    const int ix = (int)floorf(u), iy = (int)floorf(v);
    const float fx = fract_(u), fy = fract_(v);
    if (filter == 1) {
        const int x0 = border_index(bu, ix, im.y), y0 = border_index(bv, iy, im.z), x1 = border_index(bu, ix + 1, im.y), y1 = border_index(bv, iy + 1, im.z);
        return f4lerp(f4lerp(image_pixel(sc, im, x0, y0), image_pixel(sc, im, x1, y0), fx), f4lerp(image_pixel(sc, im, x0, y1), image_pixel(sc, im, x1, y1), fx), fy);
    }
    const float g0x = cubic_g0(fx), g0y = cubic_g0(fy), g1x = cubic_g1(fx), g1y = cubic_g1(fy);
    const int ix0 = (int)floorf((float)ix + cubic_h0(fx) + 0.5f), iy0 = (int)floorf((float)iy + cubic_h0(fy) + 0.5f);
    const int ix1 = (int)floorf((float)ix + cubic_h1(fx) + 0.5f), iy1 = (int)floorf((float)iy + cubic_h1(fy) + 0.5f);
    const int x0 = border_index(bu, ix0, im.y), y0 = border_index(bv, iy0, im.z), x1 = border_index(bu, ix1, im.y), y1 = border_index(bv, iy1, im.z);
    const float4 p00 = f4mulf(image_pixel(sc, im, x0, y0), g0x * g0y), p10 = f4mulf(image_pixel(sc, im, x1, y0), g1x * g0y);
    const float4 p01 = f4mulf(image_pixel(sc, im, x0, y1), g0x * g1y), p11 = f4mulf(image_pixel(sc, im, x1, y1), g1x * g1y);
    return f4add(f4add(p00, p10), f4add(p01, p11));
}

// texture `id` at ctx.uvw.xy = (u, v); a texture is 24 words: type, image, filter, border_u, border_v, 3 reserved, transform[6], p[10]
__device__ __forceinline__ C3 eval_texture(const DevScene& sc, int id, float u, float v) {
    const float* T = sc.textures + 24 * id;
    const int type = __float_as_int(__ldg(T));
    const float u2 = dot(v3(__ldg(T + 8), __ldg(T + 9), __ldg(T + 10)), v3(u, v, 1));     // mat3x3_transform_point_affine, core/matrix.art:237-240
    const float v2 = dot(v3(__ldg(T + 11), __ldg(T + 12), __ldg(T + 13)), v3(u, v, 1));
    if (type == 0) {
        const float sx = u2 * __ldg(T + 14), sy = v2 * __ldg(T + 15);
        const bool px = ((int)wrap_(sx, 0, 2) % 2) == 0, py = ((int)wrap_(sy, 0, 2) % 2) == 0;
        return (px != py) ? c3(__ldg(T + 16), __ldg(T + 17), __ldg(T + 18)) : c3(__ldg(T + 19), __ldg(T + 20), __ldg(T + 21));
    }
    const float4 c = image_filter(sc, __float_as_int(__ldg(T + 1)), __float_as_int(__ldg(T + 2)), __float_as_int(__ldg(T + 3)), __float_as_int(__ldg(T + 4)), u2, v2);
    return c3(c.x, c.y, c.z);
}

// ---- microfacets: core/microfacet.art:159-202,370-399 -------------------------------------------------------------------------------
__device__ __forceinline__ float absolute_cos(V3 a, V3 b) { return fabsf(dot(a, b)); }
__device__ __forceinline__ V3 reflect_(V3 v, V3 n) { return mulf(n, 2 * dot(n, v)) - v; }                                   // core/vector.art:124
__device__ __forceinline__ V3 to_world(const M33& l, V3 v) { return (mulf(l.c0, v.x) + mulf(l.c1, v.y)) + mulf(l.c2, v.z); }
__device__ __forceinline__ V3 to_local(const M33& l, V3 v) { return v3(dot(l.c0, v), dot(l.c1, v), dot(l.c2, v)); }
__device__ __forceinline__ float g_1_smith(const M33& local, V3 w, float au, float av) {
    const float cosZ = dot(local.c2, w);
    if (fabsf(cosZ) <= IGB_FLT_EPS) return 0;
    const float cosX = dot(local.c0, w), cosY = dot(local.c1, w);
    const float kx = au * cosX, ky = av * cosY;
    const float a2 = kx * kx + ky * ky;
    if (a2 <= IGB_FLT_EPS) return 1;
    const float k2 = a2 / (cosZ * cosZ);
    const float denom = 1 + sqrtf(1 + k2);
    return 2 / denom;
}
__device__ __forceinline__ float ndf_ggx(const M33& local, V3 m, float au, float av) {
    const float cosZ = dot(local.c2, m), cosX = dot(local.c0, m), cosY = dot(local.c1, m);
    const float kx = cosX / au, ky = cosY / av;
    const float k = kx * kx + ky * ky + cosZ * cosZ;
    return safe_div(1, IGB_FLT_PI * au * av * k * k);
}
__device__ __forceinline__ V3 sample_vndf_ggx(Rng& rnd, const M33& local, V3 vN, float au, float av) {
    const V3 vL = to_local(local, vN);
    const V3 sL = normalize(v3(au * vL.x, av * vL.y, vL.z));
    const float u0 = rnd.next_f32(); const float u1 = rnd.next_f32();
    const float phi = 2 * IGB_FLT_PI * u0;
    const float z = (1 - u1) * (1 + sL.z) - sL.z;
    const float sinTheta = sqrtf(clampf(1 - z * z, 0, 1));
    float sn, cs; dm_sincosf(phi, &sn, &cs);
    const float x = sinTheta * cs, y = sinTheta * sn;
    const V3 h = v3(x, y, z) + vL;                    // as the reference writes it (the unstretched view vector)
    const V3 Nh = normalize(v3(h.x * au, h.y * av, h.z));
    return to_world(local, Nh);
}
__device__ __forceinline__ float pdf_vndf_ggx(const M33& local, V3 w, V3 h, float au, float av) {
    const float cosZ = absolute_cos(local.c2, w);
    return safe_div(g_1_smith(local, w, au, av) * absolute_cos(w, h) * ndf_ggx(local, h, au, av), cosZ);
}

// ---- normal / bump mapping: bsdf/map.art:39-68, core/sampling.art:118-165, core/matrix.art:124-127,261-284 ---------------------------
__device__ __forceinline__ V3 ensure_valid_reflection(V3 Ng, V3 I, V3 N) {
    const V3 R = reflect_(I, N);
    const float threshold = fminf(0.9f * dot(Ng, I), 0.01f);
    if (dot(Ng, R) >= threshold) return N;
    const float NdotNg = dot(N, Ng);
    const V3 X = normalize(N - mulf(Ng, NdotNg));
    const float Ix = dot(I, X), Iz = dot(I, Ng);
    const float Ix2 = Ix * Ix, Iz2 = Iz * Iz;
    const float a = Ix2 + Iz2;
    const float b = safe_sqrt(Ix2 * (a - threshold * threshold));
    const float c = Iz * threshold + a;
    const float fac = 0.5f / a;
    const float N1_z2 = fac * (b + c), N2_z2 = fac * (-b + c);
    const bool valid1 = (N1_z2 > 1e-5f) && (N1_z2 <= (1.0f + 1e-5f)), valid2 = (N2_z2 > 1e-5f) && (N2_z2 <= (1.0f + 1e-5f));
    float nx, ny;
    if (valid1 && valid2) {
        const float n1x = safe_sqrt(1 - N1_z2), n1y = safe_sqrt(N1_z2), n2x = safe_sqrt(1 - N2_z2), n2y = safe_sqrt(N2_z2);
        const float R1 = 2 * (n1x * Ix + n1y * Iz) * n1y - Iz, R2 = 2 * (n2x * Ix + n2y * Iz) * n2y - Iz;
        const bool valid3 = R1 >= 1e-5f, valid4 = R2 >= 1e-5f;
        const bool first = (valid3 && valid4) ? (R1 < R2) : (R1 > R2);
        nx = first ? n1x : n2x; ny = first ? n1y : n2y;
    } else if (valid1 || valid2) {
        const float Nz2 = valid1 ? N1_z2 : N2_z2;
        nx = safe_sqrt(1 - Nz2); ny = safe_sqrt(Nz2);
    } else { nx = 0; ny = 1; }
    return mulf(X, nx) + mulf(Ng, ny);
}
__device__ __forceinline__ M33 align_vectors(V3 a, V3 b) {
    const V3 axis = cross(b, a);
    const float cosA = dot(a, b);
    M33 m;
    if (cosA <= -1) { m.c0 = v3(-1, 0, 0); m.c1 = v3(0, -1, 0); m.c2 = v3(0, 0, -1); return m; }
    const float k = 1 / (1 + cosA);
    m.c0 = v3((axis.x * axis.x * k) + cosA, (axis.y * axis.x * k) - axis.z, (axis.z * axis.x * k) + axis.y);
    m.c1 = v3((axis.x * axis.y * k) + axis.z, (axis.y * axis.y * k) + cosA, (axis.z * axis.y * k) - axis.x);
    m.c2 = v3((axis.x * axis.z * k) - axis.y, (axis.y * axis.z * k) + axis.x, (axis.z * axis.z * k) + cosA);
    return m;
}
__device__ __forceinline__ M33 normal_set_frame(const M33& local, V3 face_normal, V3 ray_dir, V3 normal) {
    const V3 n = ensure_valid_reflection(face_normal, neg(ray_dir), normalize(normal));
    const M33 t = align_vectors(local.c2, n);
    M33 r; r.c0 = m33_mul(t, local.c0); r.c1 = m33_mul(t, local.c1); r.c2 = m33_mul(t, local.c2);
    return r;
}

__device__ __forceinline__ float conductor_factor_(float n, float k, float cos_i) {   // core/fresnel.art:29-36
    const float f = n * n + k * k;
    const float d1 = f * cos_i * cos_i;
    const float d2 = 2.0f * n * cos_i;
    const float R_s = safe_div(d1 - d2, d1 + d2);
    const float R_p = safe_div(f - d2 + cos_i * cos_i, f + d2 + cos_i * cos_i);
    return clampf((R_s * R_s + R_p * R_p) * 0.5f, 0, 1);
}

// ---- the general BSDF (bsdf/diffuse.art:2-12, bsdf/dielectric.art:15-37, bsdf/conductor.art:2-141) ------------------------------------
struct BsdfD {
    int type; bool entering, mirror, rough;
    M33 local;
    C3 kd;            // DIFFUSE: kd | DIELECTRIC: ks | CONDUCTOR: ks
    C3 kt;            // DIELECTRIC: kt | CONDUCTOR: eta
    C3 ck;            // CONDUCTOR: k
    float n1, n2;     // DIELECTRIC iors | rough CONDUCTOR: alpha_u, alpha_v
    __device__ __forceinline__ bool is_all_delta() const { return type == 1 || (type == 2 && !rough); }
    __device__ __forceinline__ C3 fresnel_term(float c) const { return c3(conductor_factor_(kt.r, ck.r, c), conductor_factor_(kt.g, ck.g, c), conductor_factor_(kt.b, ck.b, c)); }
    __device__ __forceinline__ C3 eval(V3 in_dir, V3 out_dir) const {
        if (type == 0) { const float pc = positive_cos(in_dir, local.c2) * IGB_FLT_INV_PI; return c3(kd.r * pc, kd.g * pc, kd.b * pc); }
        if (type == 2 && rough) {
            const V3 N = local.c2;
            const float cos_o = absolute_cos(out_dir, N), cos_i = absolute_cos(in_dir, N);
            if (cos_o <= IGB_FLT_EPS || cos_i <= IGB_FLT_EPS) return c3(0, 0, 0);
            const V3 H = normalize(in_dir + out_dir);
            const float D = ndf_ggx(local, H, n1, n2);
            const float G = g_1_smith(local, in_dir, n1, n2) * g_1_smith(local, out_dir, n1, n2);
            const C3 F = fresnel_term(absolute_cos(out_dir, H));
            const float s = D * G / (4 * cos_o);
            return c3((0.0f * (1 - F.r) + kd.r * F.r) * s, (0.0f * (1 - F.g) + kd.g * F.g) * s, (0.0f * (1 - F.b) + kd.b * F.b) * s);   // kd = black in make_rough_conductor_bsdf
        }
        return c3(0, 0, 0);
    }
    __device__ __forceinline__ float pdf(V3 in_dir, V3 out_dir) const {
        if (type == 0) return positive_cos(in_dir, local.c2) / IGB_FLT_PI;
        if (type == 2 && rough) {
            const V3 H = normalize(in_dir + out_dir);
            const float jacob = safe_div(1, 4 * absolute_cos(out_dir, H));
            return pdf_vndf_ggx(local, out_dir, H, n1, n2) * jacob;
        }
        return 0.0f;
    }
};

// ---- 1-D / 2-D cdfs over a buffer without the leading 0 (core/cdf.art:34-75,105-155, core/interval.art:7-23) ---------------------------
__device__ __forceinline__ float cdf_get(const float* d, int i) { return i == 0 ? 0.0f : __ldg(d + i - 1); }
__device__ __forceinline__ int cdf_sample_discrete(const float* d, int func_size, float u, float& pdf) {
    const int size = func_size + 1;
    int first = 0, len = size;
    while (len > 0) {
        const int half = len / 2, middle = first + half;
        if (cdf_get(d, middle) <= u) { first = middle + 1; len -= half + 1; } else len = half;
    }
    const int off = min(min(max(first - 1, 0), size - 1), func_size - 1);
    pdf = cdf_get(d, off + 1) - cdf_get(d, off);
    return off;
}
__device__ __forceinline__ int cdf_sample_continuous(const float* d, int func_size, float u, float& pos, float& pdf) {
    float dpdf;
    const int off = cdf_sample_discrete(d, func_size, u, dpdf);
    const float rem = safe_div(u - cdf_get(d, off), dpdf);
    pos = clampf(((float)off + rem) / (float)func_size, 0, 1);
    pdf = dpdf * (float)func_size;
    return off;
}
__device__ __forceinline__ int cdf_pdf_continuous(const float* d, int func_size, float x, float& pdf) {
    const int off = min(max((int)(x * (float)func_size), 0), func_size - 1);
    pdf = (cdf_get(d, off + 1) - cdf_get(d, off)) * (float)func_size;
    return off;
}

// ---- textured environment lights (light/env.art:13-21,112-167); L = the light's 32 words ----------------------------------------------
__device__ __forceinline__ V3 switch_env_up(V3 v) { return v3(v.x, v.z, v.y); }
__device__ __forceinline__ void map_env_uv(V3 dir, float& u, float& v) {
    const float theta = dm_acosf(dir.z);
    float phi = dm_atan2f(dir.y, dir.x);
    if (phi < 0) phi = phi + 2 * IGB_FLT_PI;
    const float vv = theta / IGB_FLT_PI, uu = phi / (2 * IGB_FLT_PI);
    u = fract_(uu + 0.25f); v = 1 - vv;
}
__device__ __forceinline__ M33 env_transform(const float* L) {
    M33 m;
    m.c0 = v3(__ldg(L + 5), __ldg(L + 6), __ldg(L + 7)); m.c1 = v3(__ldg(L + 8), __ldg(L + 9), __ldg(L + 10)); m.c2 = v3(__ldg(L + 11), __ldg(L + 12), __ldg(L + 13));
    return m;
}
// radiance the environment emits towards -dir and (ENV_TEXTURED) the pdf with which sample_direct produces dir
__device__ __noinline__ float4 env_textured_eval(const DevScene& sc, const float* L, int type, float dx, float dy, float dz) {
    const V3 ldir = switch_env_up(m33_mul(env_transform(L), v3(dx, dy, dz)));
    float u, v; map_env_uv(ldir, u, v);
    const C3 t = eval_texture(sc, __float_as_int(__ldg(L + 14)), u, v);
    float pdf = 1 / (4 * IGB_FLT_PI);
    if (type == 8) {
        const float sinTheta = safe_sqrt(1 - ldir.z * ldir.z);
        const float* cdf = sc.aux_data + __float_as_int(__ldg(L + 15));
        const int sx = __float_as_int(__ldg(L + 16)), sy = __float_as_int(__ldg(L + 17));
        float pdf1, pdf2;
        const int off1 = cdf_pdf_continuous(cdf, sy, v, pdf1);
        cdf_pdf_continuous(cdf + sy + (size_t)off1 * sx, sx, u, pdf2);
        pdf = safe_div(pdf1 * pdf2, sinTheta * IGB_FLT_PI * IGB_FLT_PI * 2);
    }
    return make_float4(__ldg(L + 2) * t.r, __ldg(L + 3) * t.g, __ldg(L + 4) * t.b, pdf);
}

}  // namespace igb
