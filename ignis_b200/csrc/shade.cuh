// Shading of one primary-queue record: surface element, emission, next-event estimation, BSDF sample, roulette.
#pragma once

#include <cooperative_groups.h>

#include "types.cuh"
#include "material.cuh"

namespace igb {

struct Surf { bool is_entering; V3 point, face_normal; float area, inv_area; float pu, pv; M33 local; };

// tex_coords of a triangle-mesh hit: vec2_lerp2 of the three vertices' uv (shapes/trimesh.art:27-36, core/vector.art:148-151)
__device__ __forceinline__ void trimesh_texcoords(const DevScene& sc, int shape, int prim, float u, float v, float& tu, float& tv) {
    const int4 si = __ldg(sc.shape_info + 2 * shape);
    const int tex_start = __ldg(sc.shape_info + 2 * shape + 1).x;
    const int4 idx = __ldg(reinterpret_cast<const int4*>(sc.blob + si.w + prim));
    const float2* T = reinterpret_cast<const float2*>(sc.blob) + tex_start;
    const float2 t0 = __ldg(T + idx.x), t1 = __ldg(T + idx.y), t2 = __ldg(T + idx.z);
    tu = lerp2(t0.x, t1.x, t2.x, u, v); tv = lerp2(t0.y, t1.y, t2.y, u, v);
}
struct Pdf { float value; int measure; };   // 0 solid, 1 area, 2 delta  (driver/pdf.art:16-46)
__device__ __forceinline__ float pdf_as_solid(Pdf p, float cos, float dist2) { return p.measure == 1 ? p.value * dist2 / cos : (p.measure == 2 ? 1.0f : p.value); }

__device__ __forceinline__ C3 cmul(C3 a, C3 b) { return c3(a.r * b.r, a.g * b.g, a.b * b.b); }
__device__ __forceinline__ C3 cmulf(C3 a, float f) { return c3(a.r * f, a.g * f, a.b * f); }
__device__ __forceinline__ C3 cadd(C3 a, C3 b) { return c3(a.r + b.r, a.g + b.g, a.b + b.b); }
__device__ __forceinline__ C3 handle_color(const DevScene& sc, C3 c) { return sc.clamp_value > 0 ? c3(fminf(c.r, sc.clamp_value), fminf(c.g, sc.clamp_value), fminf(c.b, sc.clamp_value)) : c; }

// core/triangle.art:12-43
__device__ __forceinline__ void make_triangle(V3 v0, V3 v1, V3 v2, V3& n, float& area) {
    const V3 e1 = v2 - v0, e2 = v0 - v1, e3 = v1 - v2;
    const float x12 = e1.z * e2.y, y12 = e1.x * e2.z, z12 = e1.y * e2.x;
    const float x23 = e2.z * e3.y, y23 = e2.x * e3.z, z23 = e2.y * e3.x;
    const V3 c12 = v3(e1.y * e2.z - x12, e1.z * e2.x - y12, e1.x * e2.y - z12);
    const V3 c23 = v3(e2.y * e3.z - x23, e2.z * e3.x - y23, e2.x * e3.y - z23);
    const V3 nn = v3(fabsf(x12) < fabsf(x23) ? c12.x : c23.x, fabsf(y12) < fabsf(y23) ? c12.y : c23.y, fabsf(z12) < fabsf(z23) ? c12.z : c23.z);
    const float l = len(nn);
    n = mulf(nn, 1 / l); area = l / 2;
}

__device__ __forceinline__ V3 f4v(float4 f) { return v3(f.x, f.y, f.z); }

// shapes/trimesh.art:14-40 (for_point = false) and :41-68 (for_point = true)
__device__ __forceinline__ void trimesh_surface(const DevScene& sc, int ent, int shape, int prim, float u, float v, bool for_point,
                                                V3 rorg, V3 rdir, float dist, Surf& s) {
    const int4 si = __ldg(sc.shape_info + 2 * shape);
    const int4 idx = __ldg(reinterpret_cast<const int4*>(sc.blob + si.w + prim));
    const float4* E = sc.ent_shade + (size_t)ent * 6;
    const float4 g0 = ldg4(E), g1 = ldg4(E + 1), g2 = ldg4(E + 2), n0 = ldg4(E + 3), n1 = ldg4(E + 4), n2 = ldg4(E + 5);
    const V3 p0 = xform_point(g0, g1, g2, f4v(ldg4(sc.blob + si.y + idx.x)));
    const V3 p1 = xform_point(g0, g1, g2, f4v(ldg4(sc.blob + si.y + idx.y)));
    const V3 p2 = xform_point(g0, g1, g2, f4v(ldg4(sc.blob + si.y + idx.z)));
    V3 fn; float area;
    make_triangle(p0, p1, p2, fn, area);
    const V3 ln = lerp2(f4v(ldg4(sc.blob + si.z + idx.x)), f4v(ldg4(sc.blob + si.z + idx.y)), f4v(ldg4(sc.blob + si.z + idx.z)), u, v);
    const V3 normal = normalize(v3(dot(v3(n0.x, n0.y, n0.z), ln), dot(v3(n1.x, n1.y, n1.z), ln), dot(v3(n2.x, n2.y, n2.z), ln)));
    s.area = area; s.inv_area = safe_div(1, area); s.pu = u; s.pv = v;
    if (for_point) {
        s.is_entering = true;
        s.point = lerp2(p0, p1, p2, u, v);
        s.face_normal = fn;
        s.local = make_orthonormal(normal);
    } else {
        const bool entering = dot(rdir, fn) <= 0;
        s.is_entering = entering;
        s.point = rorg + mulf(rdir, dist);
        s.face_normal = entering ? fn : neg(fn);
        s.local = make_orthonormal(entering ? normal : neg(normal));
    }
}

// core/sampling.art:13-21,62-69
__device__ __forceinline__ void sample_cosine_hemisphere(float u, float v, V3& dir, float& pdf) {
    const float c = safe_sqrt(v), s = safe_sqrt(1 - v);
    const float phi = 2 * IGB_FLT_PI * u;
    float sn, cs; dm_sincosf(phi, &sn, &cs);
    dir = v3(s * cs, s * sn, c); pdf = c / IGB_FLT_PI;
}
// core/warp.art:63-91
__device__ __forceinline__ V3 equal_area_square_to_sphere(float px, float py) {
    const float u = 2 * px - 1, v = 2 * py - 1;
    const float au = fabsf(u), av = fabsf(v);
    const float sd = 1 - (au + av);
    const float d = fabsf(sd);
    const float r = 1 - d;
    const float phi = (r == 0 ? 1.0f : (av - au) / r + 1) * IGB_FLT_PI / 4;
    const float cosTheta = copysignf(1 - r * r, sd);
    const float sinTheta = safe_sqrt(2 - r * r) * r;
    float sn, cs; dm_sincosf(phi, &sn, &cs);
    return v3(copysignf(cs, u) * sinTheta, copysignf(sn, v) * sinTheta, cosTheta);
}

// core/fresnel.art:7-27
__device__ __forceinline__ bool fresnel(float eta, float cos_i, float& cos_t_out, float& factor) {
    const float eta2 = cos_i < 0 ? 1 / eta : eta;
    const float cos2_t = 1 - (1 - cos_i * cos_i) * eta2 * eta2;
    if (cos2_t <= 0.0f) return false;
    const float cos_t = sqrtf(cos2_t);
    cos_t_out = cos_i < 0 ? -cos_t : cos_t;
    const float ci = fabsf(cos_i);
    const float R_s = safe_div(eta2 * ci - cos_t, eta2 * ci + cos_t);
    const float R_p = safe_div(ci - eta2 * cos_t, ci + eta2 * cos_t);
    factor = clampf((R_s * R_s + R_p * R_p) * 0.5f, 0, 1);
    return true;
}

// core/fresnel.art:29-36
__device__ __forceinline__ float conductor_factor(float n, float k, float cos_i) {
    const float f = n * n + k * k;
    const float d1 = f * cos_i * cos_i;
    const float d2 = 2.0f * n * cos_i;
    const float R_s = safe_div(d1 - d2, d1 + d2);
    const float R_p = safe_div(f - d2 + cos_i * cos_i, f + d2 + cos_i * cos_i);
    return clampf((R_s * R_s + R_p * R_p) * 0.5f, 0, 1);
}

// light/area.art:124-190: spherical rectangle (Urena et al. 2013)
struct PlaneEm { V3 origin, normal, ex, ey; float area, inv_area, width, height; };
struct SQ { V3 o, n; float x0, y0, z0, x1, y1, b0, b1, k, s; };
__device__ __forceinline__ PlaneEm load_plane(const float* L) {
    PlaneEm e;
    e.origin = v3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
    const V3 xa = v3(__ldg(L + 5), __ldg(L + 6), __ldg(L + 7)), ya = v3(__ldg(L + 8), __ldg(L + 9), __ldg(L + 10));
    e.normal = v3(__ldg(L + 11), __ldg(L + 12), __ldg(L + 13));
    e.area = __ldg(L + 14);
    e.inv_area = safe_div(1, e.area);
    e.width = len(xa); e.height = len(ya);
    e.ex = mulf(xa, 1 / e.width); e.ey = mulf(ya, 1 / e.height);
    return e;
}
__device__ __forceinline__ SQ compute_sq(const PlaneEm& e, V3 from_point) {
    const V3 dir = e.origin - from_point;
    const float x0 = dot(dir, e.ex), y0 = dot(dir, e.ey), z0_ = dot(dir, e.normal);
    const float x1 = x0 + e.width, y1 = y0 + e.height;
    const bool pos = !signbit(z0_);
    const float z0 = pos ? -z0_ : z0_;
    const float df[4] = {x0 - x1, y1 - y0, x1 - x0, y0 - y1};
    const float a[4] = {y0, x1, y1, x0};
    float nz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float nz_ = a[i] * df[i];
        nz[i] = nz_ / sqrtf((df[i] * df[i]) * (z0 * z0) + nz_ * nz_);
    }
    const float g0 = dm_acosf(clampf(-nz[0] * nz[1], -1, 1)), g1 = dm_acosf(clampf(-nz[1] * nz[2], -1, 1));
    const float g2 = dm_acosf(clampf(-nz[2] * nz[3], -1, 1)), g3 = dm_acosf(clampf(-nz[3] * nz[0], -1, 1));
    SQ q;
    q.o = from_point; q.n = pos ? neg(e.normal) : e.normal;
    q.x0 = x0; q.y0 = y0; q.z0 = z0; q.x1 = x1; q.y1 = y1;
    q.b0 = nz[0]; q.b1 = nz[2];
    q.k = 2 * IGB_FLT_PI - g2 - g3;
    q.s = g0 + g1 - q.k;
    return q;
}

struct LightSample { V3 pos, dir; C3 intensity; Pdf pdf; float cos, dist; };

// light/area.art:62-107 (shape emitter over a triangle mesh entity)
__device__ __forceinline__ void shape_emitter_sample(const DevScene& sc, int entity, float uvx, float uvy, Surf& surf, float& pdfv, float& weight) {
    const float4* E = sc.ent_shade + (size_t)entity * 6;
    const int shape = __float_as_int(ldg4(E + 3).w);
    const int count = __ldg(sc.shape_info + 2 * shape + 1).y;
    const float ux = uvx * (float)count;
    const int f = min((int)ux, count - 1);
    float u = ux - (float)f, v = uvy;
    if (u + v > 1) { u = 1 - u; v = 1 - v; }
    trimesh_surface(sc, entity, shape, f, u, v, true, v3(0, 0, 0), v3(0, 0, 0), 0, surf);
    pdfv = surf.inv_area / (float)count;
    weight = surf.area * (float)count;
}

// shapes/sphere.art:29-45 (only what the light sample needs: point and normal)
__device__ __forceinline__ void sphere_point_for_normal(const float4* E, V3 origin, float radius, V3 normal, V3& point, V3& gn) {
    const float4 g0 = ldg4(E), g1 = ldg4(E + 1), g2 = ldg4(E + 2), n0 = ldg4(E + 3), n1 = ldg4(E + 4), n2 = ldg4(E + 5);
    point = xform_point(g0, g1, g2, origin + mulf(normal, radius));
    gn = normalize(v3(dot(v3(n0.x, n0.y, n0.z), normal), dot(v3(n1.x, n1.y, n1.z), normal), dot(v3(n2.x, n2.y, n2.z), normal)));
}
// light/area.art:268-294
__device__ __forceinline__ void sphere_emitter_sample(const DevScene& sc, const float* L, float u, float v, V3 from_point, V3& to_point, V3& to_normal) {
    const float4* E = sc.ent_shade + (size_t)__float_as_int(__ldg(L + 1)) * 6;
    const V3 origin = v3(__ldg(L + 5), __ldg(L + 6), __ldg(L + 7)); const float radius = __ldg(L + 8);
    const float4 g0 = ldg4(E), g1 = ldg4(E + 1), g2 = ldg4(E + 2);
    const V3 glb_org = xform_point(g0, g1, g2, origin);
    sphere_point_for_normal(E, origin, radius, equal_area_square_to_sphere(u, v), to_point, to_normal);
    const V3 os = from_point - glb_org, ps = from_point - to_point;
    if (!(len2(ps) <= len2(os))) {
        const V3 po = glb_org - to_point;
        const V3 np = to_point + mulf(po, 2);
        const V3 norm = normalize(np - glb_org);
        const float4 n0 = ldg4(E + 3), n1 = ldg4(E + 4), n2 = ldg4(E + 5);   // rows of normal_mat
        // pointmapper.art:33: (normal_mat^T n) / |diag(normal_mat)|^2
        const V3 c0 = v3(n0.x, n1.x, n2.x), c1 = v3(n0.y, n1.y, n2.y), c2 = v3(n0.z, n1.z, n2.z);
        const V3 ln = mulf(v3(dot(c0, norm), dot(c1, norm), dot(c2, norm)), 1 / len2(v3(n0.x, n1.y, n2.z)));
        sphere_point_for_normal(E, origin, radius, ln, to_point, to_normal);
    }
}

// core/warp.art:2-22
__device__ __forceinline__ void square_to_concentric_disk(float px, float py, float& x, float& y) {
    const float a = 2 * px - 1, b = 2 * py - 1;
    if (a == 0 && b == 0) { x = 0; y = 0; return; }
    float sn, cs;
    if (a * a > b * b) { const float phi = (IGB_FLT_PI / 4) * safe_div(b, a); dm_sincosf(phi, &sn, &cs); x = cs * a; y = sn * a; }
    else { const float phi = (IGB_FLT_PI / 2) - (IGB_FLT_PI / 4) * safe_div(a, b); dm_sincosf(phi, &sn, &cs); x = cs * b; y = sn * b; }
}
__device__ __forceinline__ float uniform_cone_pdf(float cos_angle) { return safe_div(1, 2 * IGB_FLT_PI * (1 - cos_angle)); }   // core/sampling.art:106

// FULL: see shade_record
template <bool FULL>
__device__ __forceinline__ LightSample light_sample_direct(const DevScene& sc, const float* L, int type, Rng& rnd, const Surf& from) {
    LightSample o;
    if (type == 0) {          // light/env.art:84-88
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        const V3 dir = equal_area_square_to_sphere(u, v);
        const float pdf = 1 / (4 * IGB_FLT_PI);
        o.pos = from.point + mulf(dir, sc.scene_radius); o.dir = dir;
        o.intensity = cmulf(c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4)), 1 / pdf);
        o.pdf.value = pdf; o.pdf.measure = 0; o.cos = 1.0f; o.dist = sc.scene_radius;
    } else if (type == 1) {   // light/point.art:3-8
        const V3 pos = v3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
        const V3 d_ = pos - from.point;
        const float dist = len(d_);
        o.pos = pos; o.dir = mulf(d_, safe_div(1, dist));
        o.intensity = c3(__ldg(L + 5), __ldg(L + 6), __ldg(L + 7));
        o.pdf.value = 1; o.pdf.measure = 1; o.cos = 1; o.dist = dist;
    } else if (FULL && type == 9) {   // make_environment_light over a texture: light/env.art:84-88,161-167
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        const V3 dir = equal_area_square_to_sphere(u, v);
        const float pdf = 1 / (4 * IGB_FLT_PI);
        const float4 e = env_textured_eval(sc, L, 9, dir.x, dir.y, dir.z);
        o.pos = from.point + mulf(dir, sc.scene_radius); o.dir = dir;
        o.intensity = cmulf(c3(e.x, e.y, e.z), 1 / pdf);
        o.pdf.value = pdf; o.pdf.measure = 0; o.cos = 1.0f; o.dist = sc.scene_radius;
    } else if (FULL && type == 8) {   // make_environment_light_textured: light/env.art:115-126,136-139 (the sampled intensity is not scaled, as in the reference)
        const float* cdf = sc.aux_data + __float_as_int(__ldg(L + 15));
        const int sx = __float_as_int(__ldg(L + 16)), sy = __float_as_int(__ldg(L + 17));
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        float p1, pdf1, p2, pdf2;
        const int off1 = cdf_sample_continuous(cdf, sy, v, p1, pdf1);
        cdf_sample_continuous(cdf + sy + (size_t)off1 * sx, sx, u, p2, pdf2);
        const C3 intensity = eval_texture(sc, __float_as_int(__ldg(L + 14)), p2, p1);
        const float theta = (1 - p1) * IGB_FLT_PI, phi = (p2 - 0.25f) * 2 * IGB_FLT_PI;
        float st, ct, sp, cp; dm_sincosf(theta, &st, &ct); dm_sincosf(phi, &sp, &cp);
        const V3 d = v3(st * cp, st * sp, ct);
        const float sinTheta = safe_sqrt(1 - d.z * d.z);
        const float pdf_dir = safe_div(pdf1 * pdf2, sinTheta * IGB_FLT_PI * IGB_FLT_PI * 2);
        const V3 dir = to_local(env_transform(L), switch_env_up(d));
        o.pos = from.point + mulf(dir, sc.scene_radius); o.dir = dir;
        o.intensity = cmulf(intensity, 1 / pdf_dir);
        o.pdf.value = pdf_dir; o.pdf.measure = 0; o.cos = 1.0f; o.dist = sc.scene_radius;
    } else if (FULL && type == 6) {   // light/sun.art:22-26; p = direction towards the sun, cos(half angle), radiance
        const float cos_angle = __ldg(L + 5);
        const M33 frame = make_orthonormal(neg(v3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4))));
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        const float c1 = 1 - cos_angle;                                                    // sample_uniform_cone, core/sampling.art:109-116
        float px, py; square_to_concentric_disk(u, v, px, py);
        const float n2 = px * px + py * py;
        const float z = cos_angle + c1 * (1 - n2);
        const float f = sqrtf(fmaxf(0.0f, c1 * (2 - c1 * n2)));
        const V3 ndir = m33_mul(frame, v3(px * f, py * f, z));
        const float inv_pdf = 2 * IGB_FLT_PI * (1 - cos_angle);
        o.pos = v3(0, 0, 0); o.dir = neg(ndir);
        o.intensity = cmulf(c3(__ldg(L + 6), __ldg(L + 7), __ldg(L + 8)), inv_pdf);
        o.pdf.value = uniform_cone_pdf(cos_angle); o.pdf.measure = 0; o.cos = z; o.dist = __int_as_float(0x7f800000);
    } else if (FULL && type == 7) {   // light/directional.art:6; p = direction the light travels, irradiance
        const V3 dir = v3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
        o.pos = from.point + mulf(dir, -sc.scene_radius); o.dir = neg(dir);
        o.intensity = c3(__ldg(L + 5), __ldg(L + 6), __ldg(L + 7));
        o.pdf.value = 1; o.pdf.measure = 2; o.cos = 1; o.dist = sc.scene_radius;
    } else if (FULL && type == 5) {   // light/spot.art:8-44
        const V3 pos = v3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4)), sdir = v3(__ldg(L + 5), __ldg(L + 6), __ldg(L + 7));
        const float cos_cutoff = __ldg(L + 8), cos_falloff = __ldg(L + 9);
        const float blend = cos_falloff - cos_cutoff;
        const V3 d_ = pos - from.point;
        const float dist = len(d_);
        const V3 out_dir = mulf(d_, safe_div(1, dist));
        const float cos_angle = dot(neg(out_dir), sdir);
        float factor;
        if (blend <= IGB_FLT_EPS) factor = cos_angle <= cos_cutoff ? 0.0f : 1.0f;
        else { const float x = clampf((cos_angle - cos_cutoff) / blend, 0, 1); factor = x * x * (3 - 2 * x); }
        o.pos = pos; o.dir = out_dir;
        o.intensity = cmulf(c3(__ldg(L + 10), __ldg(L + 11), __ldg(L + 12)), factor);
        o.pdf.value = cos_angle > cos_cutoff ? 1.0f : 0.0f; o.pdf.measure = 1;
        o.cos = -dot(out_dir, sdir); o.dist = dist;
    } else {                  // light/area.art:12-25
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        V3 to_point, to_normal; float weight; C3 radiance;
        if (type == 2) {
            const PlaneEm e = load_plane(L);
            const SQ sq = compute_sq(e, from.point);
            const float au = fma_(u, sq.s, sq.k);
            float sn, cs; dm_sincosf(au, &sn, &cs);
            const float fu = fma_(cs, sq.b0, -sq.b1) / sn;
            const float cu = clampf(copysignf(1.0f, fu) / sqrtf(sum_of_prod(fu, fu, sq.b0, sq.b0)), -1, 1);
            const float xu = clampf(-(cu * sq.z0) / sqrtf(fma_(-cu, cu, 1.0f)), sq.x0, sq.x1);
            const float d = sqrtf(sum_of_prod(xu, xu, sq.z0, sq.z0));
            const float h0 = sq.y0 / sqrtf(sum_of_prod(d, d, sq.y0, sq.y0));
            const float h1 = sq.y1 / sqrtf(sum_of_prod(d, d, sq.y1, sq.y1));
            const float hv = fma_(v, h1 - h0, h0);
            const float hv2 = hv * hv;
            const float yv = (hv2 < 1 - 1e-6f) ? (hv * d) / sqrtf(1 - hv2) : sq.y1;
            to_point = sq.o + (mulf(e.ex, xu) + (mulf(e.ey, yv) + mulf(sq.n, sq.z0)));
            to_normal = e.normal;
            o.pdf.value = safe_div(1, sq.s); o.pdf.measure = 0;
            weight = sq.s;
            radiance = c3(__ldg(L + 23), __ldg(L + 24), __ldg(L + 25));
        } else if (FULL && type == 4) {   // sphere emitter
            sphere_emitter_sample(sc, L, u, v, from.point, to_point, to_normal);
            const float area = __ldg(L + 9);
            o.pdf.value = safe_div(1, area); o.pdf.measure = 1;
            weight = area;
            radiance = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
        } else {
            Surf to; float pdfv;
            shape_emitter_sample(sc, __float_as_int(__ldg(L + 1)), u, v, to, pdfv, weight);
            to_point = to.point; to_normal = to.face_normal;
            o.pdf.value = pdfv; o.pdf.measure = 1;
            radiance = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
        }
        const V3 d_ = to_point - from.point;
        const float dist = len(d_);
        const V3 dir = mulf(d_, safe_div(1, dist));
        o.pos = to_point; o.dir = dir;
        o.cos = dot(dir, to_normal) * (from.is_entering ? -1.0f : 1.0f);
        o.intensity = cmulf(radiance, weight);
        o.dist = dist;
    }
    return o;
}

// Where shading puts its results: the next primary queue, the shadow queue and their counters. Records are appended
// from inside divergent code with one atomic per coalesced group of lanes (cooperative_groups::coalesced_threads), at
// the point where they are computed, so that nothing has to stay live until a common exit.
struct ShadeSink {
    PrimaryQueue nq; int* next_count;
    ShadowQueue sq;  int* shadow_count;
};
__device__ __forceinline__ int coalesced_append(int* counter) {
    const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
    int base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, (int)g.size());
    return g.shfl(base, 0) + (int)g.thread_rank();
}

// ---- non-uniform light selectors (light/light_selector.art:46-110). Kept out of line: scenes with the uniform selector -- every
// BASELINE configuration -- never execute them, and the shade kernel is instruction-fetch bound.
struct LightPick { int id; float pdf; uint32_t counter; };   // id: index into [infinite lights | finite lights]

// light_cdf.bin = [x1 .. x(n-1), 1]; the leading 0 is virtual (core/cdf.art:43-48,70-73)
__device__ __forceinline__ float sel_cdf_get(const float* data, int i) { return i == 0 ? 0.0f : __ldg(data + i - 1); }
__device__ __forceinline__ int sel_cdf_sample(const float* data, int n_fin, float u, float& pdf) {
    const int size = n_fin + 1;
    int first = 0, len = size;                     // interval::binary_search, core/interval.art:7-23
    while (len > 0) {
        const int half = len / 2, middle = first + half;
        if (sel_cdf_get(data, middle) <= u) { first = middle + 1; len -= half + 1; } else len = half;
    }
    const int off = min(min(max(first - 1, 0), size - 1), n_fin - 1);
    pdf = sel_cdf_get(data, off + 1) - sel_cdf_get(data, off);
    return off;
}
// light_hierarchy.bin = codes[round_up(n, 4)] then 8 words per node {pos, +-flux, dir, id} (light/light_hierarchy.art:13-38)
struct HEntry { V3 pos, dir; float flux; int id; bool has_dir, is_leaf; };
__device__ __forceinline__ HEntry sel_h_load(const float* nodes, int id) {
    const float4 e1 = __ldg(reinterpret_cast<const float4*>(nodes) + id * 2), e2 = __ldg(reinterpret_cast<const float4*>(nodes) + id * 2 + 1);
    const int index = __float_as_int(e2.w);
    HEntry h;
    h.pos = v3(e1.x, e1.y, e1.z); h.dir = v3(e2.x, e2.y, e2.z);
    h.flux = fabsf(e1.w); h.id = index < 0 ? -index - 1 : index;
    h.has_dir = (__float_as_uint(e1.w) >> 31) == 0; h.is_leaf = index >= 0;
    return h;
}
__device__ __forceinline__ float sel_h_cost(const HEntry& e, V3 pos) {      // light_hierarchy.art:40-52
    const V3 cdir = e.pos - pos;
    const float dist2 = len2(cdir);
    const float cos_d = e.has_dir ? fabsf(dot(e.dir, normalize(cdir))) : 1.0f;
    return safe_div(e.flux * cos_d, dist2);
}
__device__ __forceinline__ float sel_h_left(const HEntry& l, const HEntry& r, V3 pos) { const float cl = sel_h_cost(l, pos), cr = sel_h_cost(r, pos); return 1 / (1 + cr / cl); }

__device__ __noinline__ LightPick selector_sample(const float* data, int kind, int n_inf, int n_fin, uint32_t seed, uint32_t counter, float px, float py, float pz) {
    Rng rnd; rnd.seed = seed; rnd.counter = counter;
    const V3 from = v3(px, py, pz);
    LightPick r; r.pdf = 1.0f; r.id = 0;
    bool finite = true; float scale = 1.0f;
    if (n_inf != 0) {   // half the samples go to the infinite lights (light_selector.art:57-75,92-108)
        const float q = rnd.next_f32();
        if (q < 0.5f) { r.id = n_inf <= 1 ? 0 : rnd.next_i32(0, n_inf - 1); r.pdf = 1 / (float)n_inf * 0.5f; finite = false; }
        else scale = 1 - 0.5f;
    }
    if (finite) {
        int id = 0; float pdf = 1.0f;
        if (kind == 1) id = sel_cdf_sample(data, n_fin, rnd.next_f32(), pdf);
        else if (n_fin > 1) {                                               // light_hierarchy.art:63-76
            const float* nodes = data + (n_fin + 3) / 4 * 4;
            HEntry entry = sel_h_load(nodes, 0);
            while (!entry.is_leaf) {
                const HEntry left = sel_h_load(nodes, entry.id), right = sel_h_load(nodes, entry.id + 1);
                const float prop = sel_h_left(left, right, from);
                const bool is_left = rnd.next_f32() < prop;
                entry = is_left ? left : right;
                pdf *= is_left ? prop : 1 - prop;
            }
            id = entry.id;
        }
        r.id = n_inf + id;
        r.pdf = n_inf != 0 ? pdf * scale : pdf;
    }
    r.counter = rnd.counter;
    return r;
}

// probability with which the selector picks this light from `from` (light_selector.art:53,71,88,106; light_hierarchy.art:78-95)
__device__ __noinline__ float selector_pdf(const float* data, int kind, int n_inf, int n_fin, int infinite, int light_id, float px, float py, float pz) {
    if (infinite) return 1 / (float)n_inf * 0.5f;
    float pdf = 1.0f;
    if (kind == 1) pdf = sel_cdf_get(data, light_id + 1) - sel_cdf_get(data, light_id);
    else if (n_fin > 1) {
        const V3 from = v3(px, py, pz);
        const float* nodes = data + (n_fin + 3) / 4 * 4;
        uint32_t code = __float_as_uint(__ldg(data + light_id));
        HEntry entry = sel_h_load(nodes, 0);
        while (!entry.is_leaf) {
            const HEntry left = sel_h_load(nodes, entry.id), right = sel_h_load(nodes, entry.id + 1);
            const float prop = sel_h_left(left, right, from);
            const bool is_left = (code & 1u) == 0;
            entry = is_left ? left : right;
            pdf *= is_left ? prop : 1 - prop;
            code >>= 1;
        }
    }
    return n_inf == 0 ? pdf : pdf * (1 - 0.5f);
}


// Hit and miss shading of primary-queue record i (gpu_hit_shade / gpu_miss_shade, driver/mapping_gpu.art:123-290;
// technique/pathtracer.art:40-228). Splats emission / environment contributions (returns how many), appends the
// shadow ray of the next-event estimate and the continuation ray, if any.
// The shade kernels are built TWICE and one is picked per launch (api.cu: scene_full / std_aovs): FULL = false has only what
// the BASELINE configurations use -- diffuse and dielectric BSDFs, environment / point / plane / mesh area lights, the uniform
// light selector -- FULL = true adds smooth conductors, sphere and spot lights, the cdf and hierarchy selectors and the standard
// AOVs. The shade phase is instruction-fetch bound: code a scene never executes still costs it when it sits between
// the instructions it does execute. The reference specialises harder -- it JIT-compiles one shader per material.
template <bool FULL>
__device__ __forceinline__ int shade_record(const DevScene& sc, const RenderParams& rp, const PrimaryQueue& q, int i, float* __restrict__ fb, const ShadeSink& sink) {
    int n_splat = 0;
    const float4 o = q.org_tmin[i], d = q.dir_tmax[i];
    const uint4 st = q.state[i];
    const float4 pc = q.contrib[i];
    const int ent = q.ent[i];
    const V3 rorg = v3(o.x, o.y, o.z), rdir = v3(d.x, d.y, d.z);
    const int ray_id = (int)st.x;
    const int sample = ray_id % rp.spi;
    const int pixel = ray_id / rp.spi;
    const int target = rp.det ? ray_id : pixel;   // where this path's radiance goes: its pixel, or its own slot (deterministic accumulation)
    const int depth = (int)(st.z & 0xFFu);
    const int iter = (int)(st.z >> 8);   // the iteration that generated the path: it may be shaded by a later launch (deferred tail)
    const float eta = __uint_as_float(st.w);
    const C3 contrib = c3(pc.x, pc.y, pc.z);
    const float inv_pdf = pc.w;
    fb += (size_t)(iter & rp.ring_mask) * (size_t)rp.ring_stride;   // frame streaming: the slot of the iteration that generated the path
    const int n_lights = sc.n_inf + sc.n_fin;
    const float pdf_lights = n_lights == 0 ? 1.0f : 1 / (float)n_lights;          // light_selector.art:26-29
    const bool nee = sc.nee != 0;
    if (ent < 0) {
        // ---- on_miss, pathtracer.art:141-168
        int inflights = 0; C3 color = c3(0, 0, 0);
        for (int l = 0; l < sc.n_inf; ++l) {
            const float* L = sc.inf_lights + 32 * l;
            C3 emit; float pdf_s;
            if (FULL && __float_as_int(__ldg(L)) >= 8) {                              // textured environment: light/env.art:145-152,161-167
                const float4 e = env_textured_eval(sc, L, __float_as_int(__ldg(L)), rdir.x, rdir.y, rdir.z);
                emit = c3(e.x, e.y, e.z); pdf_s = e.w;
            } else if (FULL && __float_as_int(__ldg(L)) != 0) {
                if (__float_as_int(__ldg(L)) == 7) continue;                          // delta lights are not seen by rays (pathtracer.art:149)
                const bool hit = dot(v3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4)), rdir) >= __ldg(L + 5);   // sun.art:18,33-45
                emit = hit ? c3(__ldg(L + 6), __ldg(L + 7), __ldg(L + 8)) : c3(0, 0, 0);
                pdf_s = hit ? uniform_cone_pdf(__ldg(L + 5)) : 0.0f;
            } else {
                emit = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));                 // env.art:96
                pdf_s = 1 / (4 * IGB_FLT_PI);                                        // env.art:97
            }
            ++inflights;
            const float sel_pdf = (!FULL || sc.selector == 0) ? pdf_lights : selector_pdf(sc.selector_data, sc.selector, sc.n_inf, sc.n_fin, 1, l, rorg.x, rorg.y, rorg.z);
            const float mis = nee ? 1 / (1 + inv_pdf * sel_pdf * pdf_s) : 1.0f;
            color = cadd(color, handle_color(sc, cmulf(cmul(contrib, emit), mis)));
        }
        if (inflights > 0) { splat(fb, target, color, rp.inv_spi); ++n_splat; }
    } else {
        const float4 hh = q.hit[i];
        const int prim = __float_as_int(hh.w);
        const float dist = hh.x;
        const float4* E = sc.ent_shade + (size_t)ent * 6;
        const int shape = __float_as_int(ldg4(E + 3).w);
        const int mat_id = __float_as_int(ldg4(E + 4).w);
        const int4 si = __ldg(sc.shape_info + 2 * shape);
        Surf surf;
        if (si.x == 0) trimesh_surface(sc, ent, shape, prim, hh.y, hh.z, false, rorg, rdir, dist, surf);
        else {  // shapes/sphere.art:52-76
            const float4 g0 = ldg4(E), g1 = ldg4(E + 1), g2 = ldg4(E + 2);
            const float4 sph = ldg4(sc.blob + si.y);
            const V3 point = rorg + mulf(rdir, dist);
            const V3 dd = point - xform_point(g0, g1, g2, v3(sph.x, sph.y, sph.z));
            const float l = len(dd);
            const V3 normal = mulf(dd, 1 / l);
            surf.is_entering = true; surf.point = point; surf.face_normal = normal; surf.area = 0; surf.inv_area = 0;
            surf.pu = hh.y; surf.pv = hh.z; surf.local = make_orthonormal(normal);
        }
        const float4* M = sc.materials + 8 * mat_id;
        const float4 m0 = ldg4(M), m1 = ldg4(M + 1), m2 = ldg4(M + 2);
        const int bsdf = __float_as_int(m0.x);
        const int light_id = __float_as_int(m0.y);
        const V3 N = surf.local.c2;
        // The BSDF shader of the material (HitShader.cpp:16-53): colour parameters that are textures are looked up at ctx.uvw = tex_coords,
        // a bump / normal map replaces the frame the BSDF is built on (bsdf/map.art:39-68); emission keeps the surface's own frame.
        C3 kd = c3(m0.z, m0.w, m1.x);                                                   // DIFFUSE reflectance
        C3 d_ks = c3(m1.x, m1.y, m1.z), d_kt = c3(m1.w, m2.x, m2.y);                    // DIELECTRIC
        C3 c_ks = c3(m2.x, m2.y, m2.z);                                                 // CONDUCTOR
        M33 bl = surf.local;
        bool rough = false; float au = 0, av = 0;
        if (FULL) {
            const float4 m4 = ldg4(M + 4), m5 = ldg4(M + 5);
            const int tex0 = __float_as_int(m4.x), tex1 = __float_as_int(m4.y), map_kind = __float_as_int(m5.y);
            if (tex0 >= 0 || tex1 >= 0 || map_kind != 0) {
                float tu = surf.pu, tv = surf.pv;                                       // sphere: tex_coords = prim_coords (shapes/sphere.art:70)
                if (si.x == 0) trimesh_texcoords(sc, shape, prim, hh.y, hh.z, tu, tv);
                if (map_kind == 1) {          // make_bumpmap, texture_dx / texture_dy (texture/common.art:28-38): forward differences, delta = 0.001
                    const int mt = __float_as_int(m5.z);
                    const float delta = 0.001f;
                    const float c0 = eval_texture(sc, mt, tu, tv).r;
                    const float dx = (eval_texture(sc, mt, tu + delta, tv).r - c0) * (1 / delta);
                    const float dy = (eval_texture(sc, mt, tu, tv + delta).r - c0) * (1 / delta);
                    const V3 nn = normalize(surf.local.c2 - mulf(mulf(surf.local.c0, dx) + mulf(surf.local.c1, dy), m5.w));
                    bl = normal_set_frame(surf.local, surf.face_normal, rdir, nn);
                } else if (map_kind == 2) {   // make_normalmap, bsdf/map.art:56-60
                    const C3 c = eval_texture(sc, __float_as_int(m5.z), tu, tv);
                    const V3 oN = to_local(surf.local, normalize(v3(2 * c.r - 1, 2 * c.g - 1, 2 * c.b - 1)));
                    const V3 nn = m5.w != 1 ? normalize(surf.local.c2 + mulf(oN - surf.local.c2, m5.w)) : oN;
                    bl = normal_set_frame(surf.local, surf.face_normal, rdir, nn);
                }
                if (tex0 >= 0) { const C3 t = eval_texture(sc, tex0, tu, tv); kd = t; d_ks = t; c_ks = t; }
                if (tex1 >= 0) d_kt = eval_texture(sc, tex1, tu, tv);
            }
            au = m4.w; av = m5.x;
            rough = bsdf == 2 && __float_as_int(m4.z) == 1 && !(au <= 1e-4f || av <= 1e-4f);   // core/microfacet.art:297,403-425
        }
        const V3 bN = bl.c2;
        Rng rnd; rnd.seed = random_seed(sample, iter, rp.frame, pixel % rp.width, pixel / rp.width, rp.seed); rnd.counter = st.y;

        // ---- wrap_infobuffer_renderer, technique/internal/infobuffer.art:9-24: Normals / Albedo of the first hit, iteration 0 only
        if (FULL && rp.aov_normals && depth == 1 && iter == 0) {
            C3 albedo;
            if (bsdf == 0) albedo = kd;                                                                         // diffuse.art:10 (kd)
            else if (bsdf == 1) albedo = c3(lerp1(d_ks.r, d_kt.r, 0.5f), lerp1(d_ks.g, d_kt.g, 0.5f), lerp1(d_ks.b, d_kt.b, 0.5f));   // dielectric.art:35 color_lerp(ks, kt, 0.5)
            else if (rough) {                                                                                   // conductor.art:50-56 (kd = black)
                const float ci = absolute_cos(neg(rdir), bN);
                const C3 F = c3(conductor_factor(m0.z, m1.y, ci), conductor_factor(m0.w, m1.z, ci), conductor_factor(m1.x, m1.w, ci));
                albedo = c3(0.0f * (1 - F.r) + c_ks.r * F.r, 0.0f * (1 - F.g) + c_ks.g * F.g, 0.0f * (1 - F.b) + c_ks.b * F.b);
            }
            else if (m2.w != 0.0f) albedo = c_ks;                                                               // conductor.art:9 (ks)
            else { const float ci = dot(neg(rdir), bN); albedo = cmul(c_ks, c3(conductor_factor(m0.z, m1.y, ci), conductor_factor(m0.w, m1.z, ci), conductor_factor(m1.x, m1.w, ci))); }   // conductor.art:28-38
            splat(rp.aov_normals, target, c3(N.x, N.y, N.z), rp.inv_spi);
            splat(rp.aov_albedo, target, c3(fminf(albedo.r, 1.0f), fminf(albedo.g, 1.0f), fminf(albedo.b, 1.0f)), rp.inv_spi);   // color_saturate(albedo, 1)
        }
        // ---- on_hit, pathtracer.art:119-139
        if (light_id >= 0 && surf.is_entering) {
            const float dt = -dot(rdir, N);
            if (dt > IGB_FLT_EPS) {
                const float* L = sc.fin_lights + 32 * light_id;
                const int lt = __float_as_int(__ldg(L));
                C3 intensity; Pdf pdf;
                if (lt == 2) {
                    intensity = c3(__ldg(L + 23), __ldg(L + 24), __ldg(L + 25));
                    const PlaneEm e = load_plane(L);
                    const SQ sqv = compute_sq(e, rorg);
                    pdf.value = safe_div(1, sqv.s); pdf.measure = 0;
                } else if (FULL && lt == 4) {   // light/area.art:301-303
                    intensity = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
                    pdf.value = safe_div(1, __ldg(L + 9)); pdf.measure = 1;
                } else {
                    intensity = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
                    Surf es; float pdfv, w;
                    shape_emitter_sample(sc, __float_as_int(__ldg(L + 1)), surf.pu, surf.pv, es, pdfv, w);
                    pdf.value = pdfv; pdf.measure = 1;
                }
                const float pdf_s = pdf_as_solid(pdf, dt, dist * dist);
                const float sel_pdf = (!FULL || sc.selector == 0) ? pdf_lights : selector_pdf(sc.selector_data, sc.selector, sc.n_inf, sc.n_fin, 0, light_id, rorg.x, rorg.y, rorg.z);
                const float mis = nee ? 1 / (1 + inv_pdf * sel_pdf * pdf_s) : 1.0f;
                splat(fb, target, handle_color(sc, cmulf(cmul(contrib, intensity), mis)), rp.inv_spi); ++n_splat;
            }
        }
        const V3 out_dir = neg(rdir);
        BsdfD rb;                                    // the rough conductor (the other BSDFs are written out below)
        if (FULL && rough) {
            rb.type = 2; rb.rough = true; rb.mirror = false; rb.entering = surf.is_entering; rb.local = bl; rb.kd = c_ks;
            rb.kt = c3(m0.z, m0.w, m1.x); rb.ck = c3(m1.y, m1.z, m1.w); rb.n1 = au; rb.n2 = av;
        }
        // ---- on_shadow, pathtracer.art:52-117
        if (nee && (bsdf == 0 || (FULL && rough)) && n_lights != 0 && !(depth + 1 > sc.max_depth)) {
            int id; float light_select_pdf = pdf_lights;
            if (!FULL || sc.selector == 0) id = n_lights <= 1 ? 0 : rnd.next_i32(0, n_lights - 1);   // light_selector.art:18-24
            else {
                const LightPick pk = selector_sample(sc.selector_data, sc.selector, sc.n_inf, sc.n_fin, rnd.seed, rnd.counter, surf.point.x, surf.point.y, surf.point.z);
                id = pk.id; light_select_pdf = pk.pdf; rnd.counter = pk.counter;
            }
            const float* L = id < sc.n_inf ? sc.inf_lights + 32 * id : sc.fin_lights + 32 * (id - sc.n_inf);
            const int lt = __float_as_int(__ldg(L));
            const LightSample ls = light_sample_direct<FULL>(sc, L, lt, rnd, surf);
            const float pdf_l_s = pdf_as_solid(ls.pdf, ls.cos, ls.dist * ls.dist) * light_select_pdf;
            if (!(pdf_l_s <= IGB_FLT_EPS) && ls.cos > IGB_FLT_EPS) {
                float mis;
                if (lt == 1 || (FULL && (lt == 5 || lt == 7))) mis = 1.0f;   // delta lights
                else { const float pdf_e_s = (FULL && rough) ? rb.pdf(ls.dir, out_dir) : positive_cos(ls.dir, bN) / IGB_FLT_PI; mis = 1 / (1 + pdf_e_s / pdf_l_s); }
                const float factor = ls.pdf.value / pdf_l_s;
                const C3 ev = (FULL && rough) ? rb.eval(ls.dir, out_dir) : cmulf(kd, positive_cos(ls.dir, bN) * IGB_FLT_INV_PI);     // diffuse.art:3
                const C3 cc = handle_color(sc, cmulf(cmul(ls.intensity, cmul(contrib, ev)), mis * factor));
                if (!((cc.r + cc.g + cc.b) / 3 <= IGB_FLT_EPS)) {
                    V3 s_dir; float s_tmax;
                    if (lt == 0 || (FULL && (lt >= 6))) { s_dir = ls.dir; s_tmax = IGB_FLT_MAX; }   // infinite lights (env, sun, directional, textured env)
                    else { s_dir = ls.pos - surf.point; s_tmax = 1 - 0.001f; }
                    const int ss = coalesced_append(sink.shadow_count);
                    sink.sq.org_tmin[ss] = make_float4(surf.point.x, surf.point.y, surf.point.z, 0.001f);
                    sink.sq.dir_tmax[ss] = make_float4(s_dir.x, s_dir.y, s_dir.z, s_tmax);
                    sink.sq.color_pix[ss] = make_float4(cc.r, cc.g, cc.b, __int_as_float(target | ((iter & rp.ring_mask) << 24)));   // ring_mask != 0 only if W * H < 2^24
                }
            }
        }
        // ---- on_bounce, pathtracer.art:170-210
        if (!(depth + 1 > sc.max_depth)) {
            V3 in_dir; float s_pdf, s_eta; C3 s_color; bool is_delta;
            if (bsdf == 0) {  // diffuse.art:5-9
                const float u = rnd.next_f32(); const float v = rnd.next_f32();
                V3 ld;
                sample_cosine_hemisphere(u, v, ld, s_pdf);
                in_dir = m33_mul(bl, ld); s_color = kd; s_eta = 1; is_delta = false;
            } else if (FULL && rough) {       // conductor.art:101-122: VNDF sample, mirror about the visible normal; a rejected sample ends the path
                s_pdf = 0; s_eta = 1; is_delta = false; in_dir = v3(0, 0, 1); s_color = c3(0, 0, 0);
                if (!(absolute_cos(out_dir, bN) <= IGB_FLT_EPS)) {
                    const V3 m = sample_vndf_ggx(rnd, bl, out_dir, au, av);
                    const float m_pdf = pdf_vndf_ggx(bl, out_dir, m, au, av);
                    if (!(len2(m) <= IGB_FLT_EPS)) {
                        const V3 oH = normalize(m);
                        const V3 H = signbit(dot(oH, out_dir)) ? neg(oH) : oH;
                        in_dir = reflect_(out_dir, H);
                        if (!(absolute_cos(in_dir, bN) <= IGB_FLT_EPS)) {
                            const float jacob = 1 / (4 * absolute_cos(out_dir, H));
                            s_pdf = m_pdf * jacob;
                            s_color = cmulf(rb.eval(in_dir, out_dir), safe_div(1, s_pdf));
                        }
                    }
                }
            } else if (FULL && bsdf == 2) {   // conductor.art:2-27: mirror / smooth conductor; p = eta rgb, k rgb, ks rgb, mirror flag
                in_dir = mulf(bN, 2 * dot(bN, out_dir)) - out_dir;                                                                 // vector.art:124
                if (m2.w != 0.0f) s_color = c_ks;
                else {
                    const float cos_i = dot(out_dir, bN);
                    s_color = cmul(c_ks, c3(conductor_factor(m0.z, m1.y, cos_i), conductor_factor(m0.w, m1.z, cos_i), conductor_factor(m1.x, m1.w, cos_i)));
                }
                s_eta = 1; s_pdf = 1; is_delta = true;
            } else {          // dielectric.art:18-34
                const float n1 = m0.z, n2 = m0.w;
                const float k = surf.is_entering ? n1 / n2 : n2 / n1;
                const float cos_o = dot(out_dir, bN);
                float cos_t = 0, factor = 1;
                if (!fresnel(k, cos_o, cos_t, factor)) { cos_t = 0; factor = 1; }
                if (rnd.next_f32() > factor) { in_dir = mulf(bN, k * cos_o - cos_t) - mulf(out_dir, k); s_color = d_kt; s_eta = k; }   // vector.art:127
                else { in_dir = mulf(bN, 2 * dot(bN, out_dir)) - out_dir; s_color = d_ks; s_eta = 1; }                               // vector.art:124
                s_pdf = 1; is_delta = true;
            }
            if (!(s_pdf <= IGB_FLT_EPS)) {
                const C3 nc = cmul(contrib, s_color);
                const C3 sc2 = cmulf(nc, eta * eta);
                const float rr = (depth + 1 > sc.min_depth) ? clampf(fmaxf(fmaxf(sc2.r, sc2.g), sc2.b), 0.05f, 0.95f) : 1.0f;
                if (!(rnd.next_f32() >= rr)) {
                    const C3 fc = cmulf(nc, 1 / rr);
                    const int bs = coalesced_append(sink.next_count);
                    sink.nq.org_tmin[bs] = make_float4(surf.point.x, surf.point.y, surf.point.z, 0.001f);
                    sink.nq.dir_tmax[bs] = make_float4(in_dir.x, in_dir.y, in_dir.z, IGB_FLT_MAX);
                    sink.nq.state[bs] = make_uint4(st.x, rnd.counter, st.z + 1u, __float_as_uint(eta * s_eta));
                    sink.nq.contrib[bs] = make_float4(fc.r, fc.g, fc.b, is_delta ? 0.0f : 1 / s_pdf);
                    sink.nq.ent[bs] = (int)RAY_BOUNCE;
                }
            }
        }
    }
    return n_splat;
}

}  // namespace igb
