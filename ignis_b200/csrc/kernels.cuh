// sm_100a kernels of the wavefront path tracer: generate -> traverse (closest) -> shade -> traverse (any hit) + splat.
//
// Reference behaviour implemented (paths relative to the reference's src/artic):
//   generate        driver/mapping_gpu.art:616-669 (gpu_generate_rays), driver/emitter.art:6-31, camera/perspective.art:29-42
//   traverse        traversal/mapping_gpu.art:67-219 / traversal/mapping_cpu.art:282-518, traversal/intersection.art:74-106,223-234
//   shade           driver/mapping_gpu.art:123-214 (gpu_hit_shade), :237-290 (gpu_miss_shade), technique/pathtracer.art:40-228,
//                   shapes/trimesh.art:14-40, shapes/sphere.art:52-76, bsdf/diffuse.art:2-12, bsdf/dielectric.art:15-37,
//                   light/{area,env,point,light_selector}.art
//   shadow + splat  driver/mapping_gpu.art:79-121 (gpu_traverse_secondary with fused framebuffer add)
//
// B200 design: rays live in HBM as structure-of-float4-arrays queues so that a warp's load of one
// field is four fully used 128-byte lines; survivors are appended to the next queue with one atomic per warp
// (ballot + popc prefix), which replaces the reference's sort/compact kernels and their host round trips; BVH8
// nodes are two 128-byte lines read through the read-only path; tensor cores are not used (no dense contraction).
#pragma once

#include "device_math.cuh"

namespace igb {

struct DevScene {
    const float4* nodes;      // BVH8 nodes, 16 float4 each (top level first, then one tree per shape)
    const float4* tris;       // 4 float4 per primitive slot: (v0,n.x) (e1,n.y) (e2,n.z) (prim_id,0,0,0)
    const float4* ent_leaf;   // 8 float4 per entity: see upload in api.cu
    const float4* ent_shade;  // 6 float4 per entity: global rows 0-2, normal rows 0-2 (.w = shape_id, mat_id, 0)
    const float4* blob;       // `shapes` dyn-table data
    const int4*   shape_info; // 2 int4 per shape: (type, v_start, n_start, i_start) (tex_start_f2, n_face, 0, 0)
    const float4* materials;  // 4 float4 per material (igb200_material)
    const float*  inf_lights; // 32 words per light (igb200_light)
    const float*  fin_lights;
    int   n_ent, n_mat, n_inf, n_fin;
    float scene_radius;
    int   max_depth, min_depth;
    float clamp_value;
    int   nee;
    float eye[3], view[9];    // view = columns right, up, dir
    float scale_x, scale_y, cam_tmin, cam_tmax;
};

struct RenderParams {
    int   spi, iter, frame, seed, width, height;
    float inv_spi;
    int   tile_w, tile_h, tiles_x, rank, world;
};

struct PrimaryQueue {
    float4* org_tmin;  // org.xyz, tmin
    float4* dir_tmax;  // dir.xyz, tmax
    uint4*  state;     // ray id, rnd counter, depth, eta bits
    float4* contrib;   // contrib rgb, inv_pdf
    float4* hit;       // t, u, v, prim_id bits
    int*    ent;       // before traversal: ray flags; after: entity id (-1 = miss)
};
struct ShadowQueue {
    float4* org_tmin;
    float4* dir_tmax;
    float4* color_pix;  // colour rgb, pixel index bits
};

constexpr uint32_t RAY_CAMERA = 1, RAY_BOUNCE = 4, RAY_SHADOW = 8, RAY_TYPE_MASK = 15;
constexpr int STACK_SIZE = 96;

struct Ray { V3 org, dir; float tmin, tmax; };
struct HitR { float t, u, v; int prim, ent; };

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// Candidate ordering: nearer wins; exactly equal distance -> larger (entity, primitive) id. See DESIGN.md "Ties".
__device__ __forceinline__ bool better(float t, int ent, int prim, const HitR& h) {
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (h.prim < 0) return true;
    return ent > h.ent || (ent == h.ent && prim > h.prim);
}

// traversal/intersection.art:74-106 with the precomputed triangle of runtime/bvh/TriBVHAdapter.h:40-61
__device__ __forceinline__ bool intersect_tri(const V3 org, const V3 dir, float tmin, float tmax, float4 a, float4 b, float4 c,
                                              float& ot, float& ou, float& ov) {
    const V3 v0 = v3(a.x, a.y, a.z), e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z), n = v3(a.w, b.w, c.w);
    const V3 cc = v0 - org;
    const V3 r = cross(cc, dir);
    const float det = dot(n, dir);
    const float abs_det = fabsf(det);
    const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
    const float u = __uint_as_float(__float_as_uint(dot(r, e1)) ^ sgn);
    const float v = __uint_as_float(__float_as_uint(dot(r, e2)) ^ sgn);
    if (!(u >= 0 && v >= 0 && u + v <= abs_det && det != 0)) return false;
    const float t = __uint_as_float(__float_as_uint(dot(cc, n)) ^ sgn);
    if (!(t >= abs_det * tmin && t <= abs_det * tmax)) return false;
    const float rcp = 1 / abs_det;
    ot = t * rcp; ou = fmaxf(u * rcp, 0.0f); ov = fmaxf(v * rcp, 0.0f);
    return true;
}

// shapes/sphere.art:1-6
__device__ __forceinline__ void sphere_map_uv(V3 dir, float& u, float& v) {
    const V3 d = v3(dir.y, -dir.x, dir.z);
    const float theta = dm_acosf(d.z);
    float phi = dm_atan2f(d.y, d.x);
    if (phi < 0) phi = phi + 2 * IGB_FLT_PI;
    u = phi / (2 * IGB_FLT_PI); v = theta / IGB_FLT_PI;
}
// shapes/sphere.art:108-136
__device__ __forceinline__ bool intersect_sphere(V3 origin, float radius, V3 org, V3 dir, float rtmin, float rtmax, float& ot, float& ou, float& ov) {
    const V3 L = org - origin;
    const float S = -dot(L, dir);
    const float D2 = len2(dir);
    const float L2 = len2(L);
    const float R2 = radius * radius * D2;
    const float M2 = L2 * D2 - S * S;
    if ((S < 0) || (M2 > R2)) return false;
    const float Q = sqrtf(R2 - M2);
    const float t0_ = (S - Q) / D2, t1_ = (S + Q) / D2;
    const float t0 = t0_ > t1_ ? t1_ : t0_, t1 = t0_ > t1_ ? t0_ : t1_;
    const float tmin = t0 < rtmin ? t1 : t0;
    if (tmin >= rtmin && tmin <= rtmax) {
        const V3 d = mulf(L + mulf(dir, tmin), 1 / radius);
        ot = tmin; sphere_map_uv(d, ou, ov);
        return true;
    }
    return false;
}

// Slab test of one BVH8 child lane (traversal/intersection.art:223-234, one fma per slab).
__device__ __forceinline__ bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, V3 idir, V3 iorg, float tmin, float tmax, float& entry) {
    const float t0x = fma_(idir.x, lox, iorg.x), t1x = fma_(idir.x, hix, iorg.x);
    const float t0y = fma_(idir.y, loy, iorg.y), t1y = fma_(idir.y, hiy, iorg.y);
    const float t0z = fma_(idir.z, loz, iorg.z), t1z = fma_(idir.z, hiz, iorg.z);
    entry = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    const float exit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
    return entry <= exit;
}

// Tests the eight children of `node` and pushes the ones that are hit; the nearest ends on top of the stack.
__device__ __forceinline__ void expand_node(const float4* __restrict__ node, V3 idir, V3 iorg, float tmin, float tmax,
                                            int* stk, float* stt, int& sp) {
    const int sp0 = sp;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int4 ch = __ldg(reinterpret_cast<const int4*>(node + 12 + g));
        if (ch.x == 0) break;  // children are packed to the front
        const float4 lox = ldg4(node + 0 + g), hix = ldg4(node + 2 + g);
        const float4 loy = ldg4(node + 4 + g), hiy = ldg4(node + 6 + g);
        const float4 loz = ldg4(node + 8 + g), hiz = ldg4(node + 10 + g);
        float e;
        if (sp + 4 > STACK_SIZE) break;  // never reached by the builder's trees (depth-bounded); guards the local array
        if (slab(lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, idir, iorg, tmin, tmax, e)) { stk[sp] = ch.x; stt[sp] = e; ++sp; }
        if (ch.y != 0 && slab(lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, idir, iorg, tmin, tmax, e)) { stk[sp] = ch.y; stt[sp] = e; ++sp; }
        if (ch.z != 0 && slab(lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, idir, iorg, tmin, tmax, e)) { stk[sp] = ch.z; stt[sp] = e; ++sp; }
        if (ch.w != 0 && slab(lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, idir, iorg, tmin, tmax, e)) { stk[sp] = ch.w; stt[sp] = e; ++sp; }
    }
    // move the nearest pushed child to the top (traversal/mapping_cpu.art:463-469 keeps the same invariant)
    if (sp - sp0 > 1) {
        int best = sp0; float bt = stt[sp0];
        for (int k = sp0 + 1; k < sp; ++k) if (stt[k] < bt) { bt = stt[k]; best = k; }
        if (best != sp - 1) {
            const int tn = stk[best]; const float tt = stt[best];
            stk[best] = stk[sp - 1]; stt[best] = stt[sp - 1];
            stk[sp - 1] = tn; stt[sp - 1] = tt;
        }
    }
}

// Two-level traversal. ANY = any-hit (shadow) mode: returns as soon as one candidate is accepted.
template <bool ANY>
__device__ __forceinline__ void trace(const DevScene& sc, const Ray& ray, uint32_t flags, HitR& hit) {
    hit.t = ray.tmax; hit.u = 0; hit.v = 0; hit.prim = -1; hit.ent = -1;
    if (sc.n_ent == 0) return;
    int stk[STACK_SIZE]; float stt[STACK_SIZE];
    int sp = 0;
    const V3 idir = v3(safe_rcp(ray.dir.x), safe_rcp(ray.dir.y), safe_rcp(ray.dir.z));   // traversal/ray.art:27-39
    const V3 iorg = neg(ray.org * idir);
    stk[0] = 1; stt[0] = ray.tmin; sp = 1;
    while (sp > 0) {
        --sp;
        const int node = stk[sp];
        if (stt[sp] > hit.t) continue;
        if (node > 0) { expand_node(sc.nodes + (size_t)(node - 1) * 16, idir, iorg, ray.tmin, hit.t, stk, stt, sp); continue; }
        const int r = -node - 1;
        const int first = r >> 2, cnt = (r & 3) + 1;
        for (int k = 0; k < cnt; ++k) {
            const float4* L = sc.ent_leaf + (size_t)(first + k) * 8;
            const float4 l0 = ldg4(L), l1 = ldg4(L + 1);
            const uint32_t eflags = __float_as_uint(l0.w);
            if ((flags & RAY_TYPE_MASK) != ((flags & eflags) & RAY_TYPE_MASK)) continue;          // ray.art:51
            float en;
            {   // intersect_ray_box_single_section, intersection.art:247-256, against the ray's own tmax
                const float t0x = fma_(idir.x, l0.x, iorg.x), t1x = fma_(idir.x, l1.x, iorg.x);
                const float t0y = fma_(idir.y, l0.y, iorg.y), t1y = fma_(idir.y, l1.y, iorg.y);
                const float t0z = fma_(idir.z, l0.z, iorg.z), t1z = fma_(idir.z, l1.z, iorg.z);
                en = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), ray.tmin));
                const float ex = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), ray.tmax));
                if (!((en <= ex) & (ex >= 0))) continue;
            }
            if (!(en <= hit.t)) continue;                                                          // mapping_cpu.art:481
            const float4 r0 = ldg4(L + 2), r1 = ldg4(L + 3), r2 = ldg4(L + 4), l5 = ldg4(L + 5);
            const V3 lorg = xform_point(r0, r1, r2, ray.org);                                      // ray.art:53-59
            const V3 ldir = xform_dir(r0, r1, r2, ray.dir);
            const int ent = __float_as_int(l5.x);
            if (__float_as_int(l1.w) == 1) {  // analytic sphere
                const float4 s = ldg4(L + 6);
                float t, u, v;
                if (intersect_sphere(v3(s.x, s.y, s.z), s.w, lorg, ldir, ray.tmin, ray.tmax, t, u, v) && better(t, ent, 0, hit)) {
                    hit.t = t; hit.u = u; hit.v = v; hit.prim = 0; hit.ent = ent;
                    if (ANY) return;
                }
                continue;
            }
            const V3 lidir = v3(safe_rcp(ldir.x), safe_rcp(ldir.y), safe_rcp(ldir.z));
            const V3 liorg = neg(lorg * lidir);
            const int base = sp;
            stk[sp] = __float_as_int(l5.y); stt[sp] = ray.tmin; ++sp;
            while (sp > base) {
                --sp;
                const int n2 = stk[sp];
                if (stt[sp] > hit.t) continue;
                if (n2 > 0) { expand_node(sc.nodes + (size_t)(n2 - 1) * 16, lidir, liorg, ray.tmin, hit.t, stk, stt, sp); continue; }
                const int rr = -n2 - 1;
                const int f2 = rr >> 2, c2 = (rr & 3) + 1;
                for (int j = 0; j < c2; ++j) {
                    const float4* T = sc.tris + (size_t)(f2 + j) * 4;
                    const float4 a = ldg4(T), b = ldg4(T + 1), c = ldg4(T + 2);
                    float t, u, v;
                    if (intersect_tri(lorg, ldir, ray.tmin, ray.tmax, a, b, c, t, u, v)) {
                        const int prim = __float_as_int(ldg4(T + 3).x);
                        if (better(t, ent, prim, hit)) {
                            hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; hit.ent = ent;
                            if (ANY) return;
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ shading helpers
struct Surf { bool is_entering; V3 point, face_normal; float area, inv_area; float pu, pv; M33 local; };
struct Pdf { float value; int measure; };   // 0 solid, 1 area, 2 delta  (driver/pdf.art:16-46)
__device__ __forceinline__ float pdf_as_solid(Pdf p, float cos, float dist2) { return p.measure == 1 ? p.value * dist2 / cos : (p.measure == 2 ? 1.0f : p.value); }

struct C3 { float r, g, b; };
__device__ __forceinline__ C3 c3(float r, float g, float b) { C3 c; c.r = r; c.g = g; c.b = b; return c; }
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

// Warp-aggregated append: one atomic per warp, compacted slot per lane (replaces gpu_compact_primary,
// driver/mapping_gpu.art:686-711, and its host read-back). Must be reached by all 32 lanes.
__device__ __forceinline__ int warp_append(int* counter, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

__device__ __forceinline__ void splat(float* fb, int pixel, C3 c, float inv_spi) {   // driver/accumulator.art:4-21
    atomicAdd(fb + (size_t)pixel * 3 + 0, c.r * inv_spi);
    atomicAdd(fb + (size_t)pixel * 3 + 1, c.g * inv_spi);
    atomicAdd(fb + (size_t)pixel * 3 + 2, c.b * inv_spi);
}

}  // namespace igb
