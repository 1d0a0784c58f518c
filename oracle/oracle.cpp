// TEST INFRASTRUCTURE -- CPU oracle for the ignis_b200 hot path. NOT part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// What it is: a scalar C++ restatement of the CPU device of PearCoding/Ignis for the `path` integrator:
// ray generation -> two-level BVH traversal -> Moeller-Trumbore -> hit/miss shading (emission, NEE, BSDF
// sample, Russian roulette) -> shadow any-hit -> framebuffer splat, following the Artic sources function by
// function (each function names the file:line it follows; paths are relative to the reference's src/artic).
//
// Pinning: checked against the reference's own known-answer tests (src/tests/artic/test_intersection.art),
// its analytic scene averages (src/tests/integrator/test_lights.py) and its converged evaluation images
// (scenes/evaluation/references, RelMSE rule of scripts/RunEvaluations.py) -- see tests/test_oracle_*.py.
// Bit-parity with the real AnyDSL-compiled CPU device: PARITY UNPINNED (the reference cannot be built here,
// it is compiled with -ffast-math and calls the platform libm).
//
// Deliberate, documented departures from a literal transcription (none changes the estimator):
//  * transcendentals come from oracle/detmath.h (see its header);
//  * the ray/box slab test uses one fused multiply-add per slab (the reference writes inv_dir*b + inv_org,
//    traversal/intersection.art:223-234, which its fast-math x86 build contracts as well);
//  * exact-distance ties between two primitives are resolved by (entity id, primitive id) instead of by BVH
//    visiting order, because the reference's order depends on an un-pinned third-party builder
//    (madmann91/bvh master, cmake/GetDependencies.cmake:53-58); for the same reason the upper distance bound
//    of the triangle test is the ray's own tmax and the running closest hit is applied afterwards, and the
//    running hit prunes nodes only through a conservative box test (cull_box) -- never through the leaf's
//    `tmin <= hit.distance` shortcut (traversal/mapping_cpu.art:481), whose outcome under rounding depends on the
//    visiting order. The closest hit is thereby a pure function of the ray: min over all primitives of
//    entities that pass the visibility and entity-box tests, ordered by (t, entity, primitive);
//  * the BVH is this file's own median-split BVH2 (or none at all: brute force), because closest-hit
//    results do not depend on topology once ties are ordered;
//  * streams are sorted with a stable counting sort instead of the in-place cycle sort
//    (driver/mapping_cpu.art:63-103): per-ray results are identical, only the order of float additions
//    into one pixel inside one iteration can differ.
//
// Build: see oracle/Makefile (-O2 -ffp-contract=off -mfma: no contraction except the explicit fmaf calls).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "detmath.h"
// The timing arm's tree is built with the product's binned-SAH builder (a header-only host function; results do not depend on the tree)
#include "../ignis_b200/csrc/bvh8.h"

namespace {

// ------------------------------------------------------------------------------------------ constants
// core/common.art:3-8
constexpr float flt_eps    = 1.1920928955e-07f;
constexpr float flt_max    = 3.4028234664e+38f;
constexpr float flt_pi     = 3.14159265359f;
constexpr float flt_inv_pi = 0.31830988618379067154f;

inline float fmaf_(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// ------------------------------------------------------------------------------------------ vectors
struct Vec2 { float x, y; };
struct Vec3 { float x, y, z; };
struct Color { float r, g, b; };  // alpha is never observable on this path

inline Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 neg(Vec3 a) { return v3(-a.x, -a.y, -a.z); }
inline Vec3 mulf(Vec3 a, float t) { return v3(a.x * t, a.y * t, a.z * t); }
// core/vector.art:98-100
inline float dot(Vec3 a, Vec3 b) { return fmaf_(a.x, b.x, fmaf_(a.y, b.y, a.z * b.z)); }
// core/vector.art:106-109
inline Vec3 cross(Vec3 a, Vec3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float len2(Vec3 a) { return dot(a, a); }
inline float len(Vec3 a) { return sqrtf(len2(a)); }
inline Vec3 normalize(Vec3 a) { return mulf(a, 1.0f / len(a)); }  // core/vector.art:140-142
// core/common.art:254-255
inline float lerp2(float a, float b, float c, float k1, float k2) { return (1 - k1 - k2) * a + k1 * b + k2 * c; }
inline Vec3 lerp2(Vec3 a, Vec3 b, Vec3 c, float u, float v) { return v3(lerp2(a.x, b.x, c.x, u, v), lerp2(a.y, b.y, c.y, u, v), lerp2(a.z, b.z, c.z, u, v)); }
inline float lerp(float a, float b, float k) { return (1 - k) * a + k * b; }

inline Color col(float r, float g, float b) { return Color{r, g, b}; }
inline Color cmul(Color a, Color b) { return col(a.r * b.r, a.g * b.g, a.b * b.b); }
inline Color cmulf(Color a, float f) { return col(a.r * f, a.g * f, a.b * f); }
inline Color cadd(Color a, Color b) { return col(a.r + b.r, a.g + b.g, a.b + b.b); }
inline float caverage(Color c) { return (c.r + c.g + c.b) / 3; }                         // core/color.art:28
inline float cmaxcomp(Color c) { return fmaxf(fmaxf(c.r, c.g), c.b); }                    // core/color.art:36, vector.art:117
inline Color csaturate(Color a, float f) { return col(fminf(a.r, f), fminf(a.g, f), fminf(a.b, f)); }  // color.art:26

// core/common.art:208-215
inline float prodsign(float x, float y) { return u2f(f2u(x) ^ (f2u(y) & 0x80000000u)); }
inline float safe_rcp(float x) { return ((x > 0 ? x : -x) < 1e-8f) ? prodsign(flt_max, x) : 1.0f / x; }
// core/common.art:285-289
inline float clampf(float v, float l, float u) { return fminf(u, fmaxf(l, v)); }
inline float safe_div(float a, float b) { return fabsf(b) <= flt_eps ? 0.0f : a / b; }
inline float safe_sqrt(float a) { return sqrtf(fmaxf(0.0f, a)); }
// core/common.art:257-272
inline float sum_of_prod(float a, float b, float c, float d) { float cd = c * d; float s = fmaf_(a, b, cd); float e = fmaf_(c, d, -cd); return s + e; }
// core/common.art:274-282
inline float positive_cos(Vec3 a, Vec3 b) { float c = dot(a, b); return c >= 0 ? c : 0.0f; }

// ------------------------------------------------------------------------------------------ matrices (column major)
struct Mat3x3 { Vec3 c0, c1, c2; };
struct Mat3x4 { Vec3 c0, c1, c2, c3; };
// core/matrix.art:105-108 (rows dotted with v, dot = nested fma)
inline Vec3 mat3x3_mul(const Mat3x3& m, Vec3 v) {
    return v3(dot(v3(m.c0.x, m.c1.x, m.c2.x), v), dot(v3(m.c0.y, m.c1.y, m.c2.y), v), dot(v3(m.c0.z, m.c1.z, m.c2.z), v));
}
// core/matrix.art:115-118, 246-247; vec4_dot = fma(a.x,b.x, fma(a.y,b.y, fma(a.z,b.z, a.w*b.w)))
inline float dot4(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw) { return fmaf_(ax, bx, fmaf_(ay, by, fmaf_(az, bz, aw * bw))); }
inline Vec3 transform_point(const Mat3x4& m, Vec3 v) {
    return v3(dot4(m.c0.x, m.c1.x, m.c2.x, m.c3.x, v.x, v.y, v.z, 1), dot4(m.c0.y, m.c1.y, m.c2.y, m.c3.y, v.x, v.y, v.z, 1), dot4(m.c0.z, m.c1.z, m.c2.z, m.c3.z, v.x, v.y, v.z, 1));
}
inline Vec3 transform_direction(const Mat3x4& m, Vec3 v) {
    return v3(dot4(m.c0.x, m.c1.x, m.c2.x, m.c3.x, v.x, v.y, v.z, 0), dot4(m.c0.y, m.c1.y, m.c2.y, m.c3.y, v.x, v.y, v.z, 0), dot4(m.c0.z, m.c1.z, m.c2.z, m.c3.z, v.x, v.y, v.z, 0));
}
// core/matrix.art:24-32 (Duff et al.)
inline Mat3x3 make_orthonormal(Vec3 n) {
    const float sign = copysignf(1.0f, n.z);
    const float a = -1 / (sign + n.z);
    const float b = n.x * n.y * a;
    Mat3x3 m;
    m.c0 = v3(1 + sign * n.x * n.x * a, sign * b, -sign * n.x);
    m.c1 = v3(b, sign + n.y * n.y * a, -n.y);
    m.c2 = n;
    return m;
}

// ------------------------------------------------------------------------------------------ RNG (core/random.art)
inline uint32_t hash_combine(uint32_t h, uint32_t d) {  // :7-13
    h = (h * 16777619u) ^ (d & 0xFF);
    h = (h * 16777619u) ^ ((d >> 8) & 0xFF);
    h = (h * 16777619u) ^ ((d >> 16) & 0xFF);
    h = (h * 16777619u) ^ ((d >> 24) & 0xFF);
    return h;
}
inline uint32_t sample_tea_u32(uint32_t v0, uint32_t v1) {  // :15-24
    uint32_t sum = 0;
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v1;
}
inline uint32_t create_random_seed(int sample, int iter, int frame, int x, int y, int user) {  // :34-43
    uint32_t h = 0x811C9DC5u;
    h = hash_combine(h, (uint32_t)sample);
    h = hash_combine(h, (uint32_t)iter);
    h = hash_combine(h, (uint32_t)frame);
    h = hash_combine(h, (uint32_t)x);
    h = hash_combine(h, (uint32_t)y);
    h = hash_combine(h, (uint32_t)user);
    return h;
}
struct Rng {  // create_random_generator_live :84-87, base :45-82
    uint32_t seed, counter;
    uint32_t next_u32() { return sample_tea_u32(seed, counter++); }
    float next_f32() { uint32_t x = next_u32(); return u2f((x & 0x7FFFFFu) | 0x3F800000u) - 1; }
    uint32_t next_u32_range(uint32_t range) {
        if (range == 0xFFFFFFFFu) return next_u32();
        const uint32_t erange = range + 1, scaling = 0xFFFFFFFFu / erange, past = erange * scaling;
        uint32_t ret = next_u32();
        while (ret >= past) ret = next_u32();
        return ret / scaling;
    }
    int next_i32(int s, int e) { return (int)next_u32_range((uint32_t)(e - s)) + s; }
};

// ------------------------------------------------------------------------------------------ ray / hit (traversal/ray.art, intersection.art)
constexpr uint32_t ray_flag_camera = 1, ray_flag_bounce = 4, ray_flag_shadow = 8, ray_flag_type_mask = 15;
struct Ray { Vec3 org, dir, inv_dir, inv_org; float tmin, tmax; uint32_t flags; };
inline Ray make_ray(Vec3 org, Vec3 dir, float tmin, float tmax, uint32_t flags) {  // ray.art:27-39
    Ray r;
    r.org = org; r.dir = dir;
    r.inv_dir = v3(safe_rcp(dir.x), safe_rcp(dir.y), safe_rcp(dir.z));
    r.inv_org = neg(org * r.inv_dir);
    r.tmin = tmin; r.tmax = tmax; r.flags = flags;
    return r;
}
inline bool check_ray_visibility(const Ray& ray, uint32_t flags) {  // ray.art:51
    return (ray.flags & ray_flag_type_mask) == ((ray.flags & flags) & ray_flag_type_mask);
}
inline Ray transform_ray(const Ray& ray, const Mat3x4& m) {  // ray.art:53-59 (direction NOT renormalised)
    return make_ray(transform_point(m, ray.org), transform_direction(m, ray.dir), ray.tmin, ray.tmax, ray.flags);
}
struct Hit { float distance; Vec2 prim_coords; int prim_id, ent_id; };
inline Hit invalid_hit(float tmax) { return Hit{tmax, Vec2{0, 0}, -1, -1}; }

struct Tri { Vec3 v0, e1, e2, n; };

// traversal/intersection.art:74-106 (no culling on this path: shapes/trimesh.art:133)
inline bool intersect_ray_tri_mt(bool backface_culling, const Ray& ray, const Tri& tri, float& ot, float& ou, float& ov) {
    const Vec3 c = tri.v0 - ray.org;
    const Vec3 r = cross(c, ray.dir);
    const float det = dot(tri.n, ray.dir);
    const float abs_det = fabsf(det);
    const uint32_t sgn = f2u(det) & 0x80000000u;
    const float u = u2f(f2u(dot(r, tri.e1)) ^ sgn);
    const float v = u2f(f2u(dot(r, tri.e2)) ^ sgn);
    bool mask = u >= 0;
    mask &= v >= 0;
    mask &= u + v <= abs_det;
    if (backface_culling) mask &= det < 0; else mask &= det != 0;
    if (!mask) return false;
    const float t = u2f(f2u(dot(c, tri.n)) ^ sgn);
    mask &= t >= abs_det * ray.tmin;
    mask &= t <= abs_det * ray.tmax;
    if (!mask) return false;
    const float rcp = 1 / abs_det;
    ot = t * rcp; ou = fmaxf(u * rcp, 0.0f); ov = fmaxf(v * rcp, 0.0f);
    return true;
}

struct BBox { Vec3 min, max; };
inline float fmin_sel(float x, float y) { return x < y ? x : y; }  // make_default_min_max, intersection.art:47-51
inline float fmax_sel(float x, float y) { return x > y ? x : y; }
// traversal/intersection.art:223-234 (unordered form)
inline void intersect_ray_box(const Ray& ray, const BBox& b, float tmax, float& entry, float& exit) {
    const float t0x = fmaf_(ray.inv_dir.x, b.min.x, ray.inv_org.x), t1x = fmaf_(ray.inv_dir.x, b.max.x, ray.inv_org.x);
    const float t0y = fmaf_(ray.inv_dir.y, b.min.y, ray.inv_org.y), t1y = fmaf_(ray.inv_dir.y, b.max.y, ray.inv_org.y);
    const float t0z = fmaf_(ray.inv_dir.z, b.min.z, ray.inv_org.z), t1z = fmaf_(ray.inv_dir.z, b.max.z, ray.inv_org.z);
    entry = fmax_sel(fmax_sel(fmin_sel(t0x, t1x), fmin_sel(t0y, t1y)), fmax_sel(fmin_sel(t0z, t1z), ray.tmin));
    exit  = fmin_sel(fmin_sel(fmax_sel(t0x, t1x), fmax_sel(t0y, t1y)), fmin_sel(fmax_sel(t0z, t1z), tmax));
}

// Conservative node culling used by this file's own BVH (not part of the reference's arithmetic): a node is skipped
// only if the slab interval, widened by a bound of its rounding error (|inv_org| * 2^-20 per axis) misses
// [tmin, t_closest * (1 + 2^-16)]. A candidate is therefore never lost to the rounding of a box test, and the
// closest hit equals the brute-force answer ordered by (t, entity, primitive) whatever the tree looks like.
inline bool cull_box(const Ray& ray, const BBox& b, float t_closest) {
    const float k = 9.5367431640625e-07f;  // 2^-20
    float entry = ray.tmin, exit = t_closest * 1.0000152587890625f;  // 1 + 2^-16
    auto axis = [&](float inv_dir, float inv_org, float lo, float hi) {
        const float a = fabsf(inv_org);
        if (a == INFINITY) return;   // org * flt_max overflowed (axis-parallel ray): no bound from this axis
        const float t0 = fmaf_(inv_dir, lo, inv_org), t1 = fmaf_(inv_dir, hi, inv_org);
        const float tn = fmin_sel(t0, t1) - a * k, tf = fmax_sel(t0, t1) + a * k;
        if (tn > entry) entry = tn;
        if (tf < exit) exit = tf;
    };
    axis(ray.inv_dir.x, ray.inv_org.x, b.min.x, b.max.x);
    axis(ray.inv_dir.y, ray.inv_org.y, b.min.y, b.max.y);
    axis(ray.inv_dir.z, ray.inv_org.z, b.min.z, b.max.z);
    return exit < entry;
}

// shapes/sphere.art:1-6 with core/warp.art:50-54
inline Vec2 sphere_map_uv(Vec3 dir) {
    const Vec3 d = v3(dir.y, -dir.x, dir.z);
    const float theta = dm_acosf(d.z);
    float phi = dm_atan2f(d.y, d.x);
    if (phi < 0) phi = phi + 2 * flt_pi;
    return Vec2{phi / (2 * flt_pi), theta / flt_pi};
}
// shapes/sphere.art:108-136
inline bool intersect_sphere(Vec3 origin, float radius, const Ray& ray, float& ot, Vec2& ouv) {
    const Vec3 L = ray.org - origin;
    const float S = -dot(L, ray.dir);
    const float D2 = len2(ray.dir);
    const float L2 = len2(L);
    const float R2 = radius * radius * D2;
    const float M2 = L2 * D2 - S * S;
    if ((S < 0) || (M2 > R2)) return false;
    const float Q = sqrtf(R2 - M2);
    const float t0_ = (S - Q) / D2, t1_ = (S + Q) / D2;
    const float t0 = t0_ > t1_ ? t1_ : t0_, t1 = t0_ > t1_ ? t0_ : t1_;
    const float tmin = t0 < ray.tmin ? t1 : t0;
    if (tmin >= ray.tmin && tmin <= ray.tmax) {
        const Vec3 dir = mulf(L + mulf(ray.dir, tmin), 1 / radius);
        ot = tmin; ouv = sphere_map_uv(dir);
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------ scene description (boundary structs)
// Same binary layouts as include/igb200.h (declared independently on purpose).
struct LookupEntry { uint32_t type_id, flags; uint64_t offset; };
struct EntityLeaf { float min[3]; int32_t entity_id; float max[3]; int32_t shape_id; float local[12]; uint32_t flags; int32_t mat_id, user1, user2; };
struct MaterialDesc { int32_t bsdf, light_id; float p[14]; int32_t tex[2]; int32_t distribution; float alpha_u, alpha_v; int32_t map_kind, map_tex; float map_strength; int32_t reserved[8]; };
struct TextureDesc { int32_t type, image, filter, border_u, border_v, reserved[3]; float transform[6]; float p[10]; };
struct ImageDesc { int32_t format, width, height, reserved; const void* pixels; };
struct LightDesc { int32_t type, entity_id; float p[30]; };
struct CameraDesc { float eye[3], dir[3], up[3]; float fov; int32_t fov_vertical; float aspect, tmin, tmax; };
struct TechniqueDesc { int32_t max_depth, min_depth; float clamp; int32_t nee; int32_t light_selector; };   // selector: 0 uniform, 1 cdf ("simple"), 2 hierarchy
struct SceneDesc {
    const float* entities; int32_t n_entities;
    const LookupEntry* shape_lookups; int32_t n_shapes;
    const uint8_t* shape_data; uint64_t shape_data_bytes;
    const EntityLeaf* leaves; int32_t n_leaves;
    const int32_t* entity_per_material; int32_t n_materials;
    const MaterialDesc* materials;
    const LightDesc* infinite_lights; int32_t n_infinite;
    const LightDesc* finite_lights; int32_t n_finite;
    CameraDesc camera;
    TechniqueDesc technique;
    float bbox_min[3], bbox_max[3];
    const float* selector_data; int32_t n_selector_data;   // light_cdf.bin / light_hierarchy.bin as 32-bit words (LoaderLight.cpp:423-476, LightHierarchy.cpp)
    const TextureDesc* textures; int32_t n_textures;
    const ImageDesc* images; int32_t n_images;
    const float* aux_data; int32_t n_aux_data;             // 2-D cdfs of the textured environment lights (CDF.cpp:70-150)
};
struct Settings { int32_t device, thread_count, spi, frame, iter, width, height, seed; };  // driver/settings.art:2-11
struct StreamRay { float org[3], dir[3], tmin, tmax; };                                       // traversal/ray.art:2-7
struct HitRecord { int32_t ent_id, prim_id; float t, u, v; };
static_assert(sizeof(EntityLeaf) == 96 && sizeof(MaterialDesc) == 128 && sizeof(LightDesc) == 128 && sizeof(TextureDesc) == 96 && sizeof(ImageDesc) == 24, "layout");

enum { SHAPE_TRIMESH = 0, SHAPE_SPHERE = 1 };
enum { BSDF_DIFFUSE = 0, BSDF_DIELECTRIC = 1, BSDF_CONDUCTOR = 2 };
enum { LIGHT_ENV_CONST = 0, LIGHT_POINT = 1, LIGHT_PLANE_AREA = 2, LIGHT_SHAPE_AREA = 3, LIGHT_SPHERE_AREA = 4, LIGHT_SPOT = 5,
       LIGHT_SUN = 6,            // make_sun_light (light/sun.art:10-48): infinite cone light; p = direction towards the sun, cos(half angle), radiance
       LIGHT_DIRECTIONAL = 7,    // make_directional_light (light/directional.art:1-17): infinite delta light; p = direction the light travels, irradiance
       LIGHT_ENV_TEXTURED = 8,   // make_environment_light_textured (light/env.art:112-160): p = scale rgb, transform (9, column major), texture, cdf offset, size_x, size_y (int bits)
       LIGHT_ENV_TEX = 9 };      // make_environment_light with a texture (light/env.art:161-167): p = scale rgb, transform (9), texture (int bits)
enum { MICROFACET_DELTA = 0, MICROFACET_VNDF_GGX = 1 };
enum { MAP_NONE = 0, MAP_BUMP = 1, MAP_NORMAL = 2 };
enum { TEX_CHECKERBOARD = 0, TEX_IMAGE = 1 };
enum { FILTER_NEAREST = 0, FILTER_BILINEAR = 1, FILTER_BICUBIC = 2 };
enum { BORDER_REPEAT = 0, BORDER_CLAMP = 1, BORDER_MIRROR = 2 };
enum { IMAGE_RGBA8 = 0, IMAGE_MONO8 = 1, IMAGE_RGBA32F = 2 };

// ------------------------------------------------------------------------------------------ own BVH2 (median split)
struct Bvh2 {
    struct Node { BBox box; int left, right, first, count; };  // count>0: leaf over order[first..first+count)
    std::vector<Node> nodes;
    std::vector<int> order;
    void build(const std::vector<BBox>& boxes, int leaf_size) {
        const int n = (int)boxes.size();
        order.resize(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        nodes.clear();
        if (n) rec(boxes, 0, n, leaf_size);
    }
    int rec(const std::vector<BBox>& boxes, int b, int e, int leaf_size) {
        const int id = (int)nodes.size();
        nodes.push_back(Node{});
        BBox bb{v3(INFINITY, INFINITY, INFINITY), v3(-INFINITY, -INFINITY, -INFINITY)}, cb = bb;
        for (int i = b; i < e; ++i) {
            const BBox& x = boxes[order[i]];
            bb.min = v3(std::min(bb.min.x, x.min.x), std::min(bb.min.y, x.min.y), std::min(bb.min.z, x.min.z));
            bb.max = v3(std::max(bb.max.x, x.max.x), std::max(bb.max.y, x.max.y), std::max(bb.max.z, x.max.z));
            const Vec3 c = mulf(x.min + x.max, 0.5f);
            cb.min = v3(std::min(cb.min.x, c.x), std::min(cb.min.y, c.y), std::min(cb.min.z, c.z));
            cb.max = v3(std::max(cb.max.x, c.x), std::max(cb.max.y, c.y), std::max(cb.max.z, c.z));
        }
        nodes[id].box = bb;
        if (e - b <= leaf_size) { nodes[id].first = b; nodes[id].count = e - b; nodes[id].left = nodes[id].right = -1; return id; }
        const Vec3 d = cb.max - cb.min;
        const int axis = d.x >= d.y ? (d.x >= d.z ? 0 : 2) : (d.y >= d.z ? 1 : 2);
        auto key = [&](int i) { const BBox& x = boxes[i]; return axis == 0 ? x.min.x + x.max.x : axis == 1 ? x.min.y + x.max.y : x.min.z + x.max.z; };
        const int mid = (b + e) / 2;
        std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e, [&](int a, int c) { const float ka = key(a), kc = key(c); return ka < kc || (ka == kc && a < c); });
        nodes[id].count = 0; nodes[id].first = 0;
        const int l = rec(boxes, b, mid, leaf_size);
        const int r = rec(boxes, mid, e, leaf_size);
        nodes[id].left = l; nodes[id].right = r;
        return id;
    }
};

// ------------------------------------------------------------------------------------------ BVH4 + 4-primitive leaves, near child first
// What the reference's CPU device walks at vector width < 8: a SAH-built tree collapsed to arity 4 with Tri4 leaves
// (src/runtime/shape/TriMeshProvider.cpp:549-557, src/runtime/bvh/NArityBvh.h:94-143, TriBVHAdapter.h:94-141), children tested unordered and
// pushed so that the nearest ends up on top of the stack (traversal/mapping_cpu.art:353-381, stack.art push / push_after). Used by the
// TIMING arm (use_bvh = 2: bench.py cpu_baseline / --impl reference); culling stays the conservative one, so hit records are identical to
// the other two walks (tests/test_oracle_kat.py).
struct Bvh4 {
    struct Node { BBox box[4]; int child[4]; };   // child > 0: node index + 1; < 0: ~leaf index; 0: empty
    struct Leaf { int first, count; };
    std::vector<Node> nodes; std::vector<Leaf> leaves; std::vector<int> order;
    void build(const std::vector<BBox>& boxes) {
        nodes.clear(); leaves.clear(); order.clear();
        if (boxes.empty()) return;
        std::vector<igb::Box3> b3(boxes.size());
        for (size_t i = 0; i < boxes.size(); ++i) b3[i] = igb::Box3{{boxes[i].min.x, boxes[i].min.y, boxes[i].min.z}, {boxes[i].max.x, boxes[i].max.y, boxes[i].max.z}};
        igb::detail::Builder2 b2(b3, 4);
        order = b2.order;
        const auto& n2 = b2.nodes;
        auto to_box = [](const igb::Box3& b) { return BBox{v3(b.lo[0], b.lo[1], b.lo[2]), v3(b.hi[0], b.hi[1], b.hi[2])}; };
        struct Work { int n2, n4; };
        std::vector<Work> st;
        nodes.push_back(Node{});
        st.push_back(Work{0, 0});
        while (!st.empty()) {
            const Work w = st.back(); st.pop_back();
            int kids[4]; int nk = 0;
            if (n2[w.n2].count > 0) kids[nk++] = w.n2;
            else {
                kids[nk++] = n2[w.n2].left; kids[nk++] = n2[w.n2].right;
                while (nk < 4) {   // open the inner child with the largest surface area (NArityBvh.h:94-143)
                    int best = -1; float best_area = -1;
                    for (int i = 0; i < nk; ++i) if (n2[kids[i]].count == 0) { const float a = n2[kids[i]].box.half_area(); if (a > best_area) { best_area = a; best = i; } }
                    if (best < 0) break;
                    const int k = kids[best];
                    kids[best] = n2[k].left; kids[nk++] = n2[k].right;
                }
            }
            Node nd{};
            for (int i = 0; i < 4; ++i) { nd.child[i] = 0; nd.box[i] = BBox{v3(INFINITY, INFINITY, INFINITY), v3(-INFINITY, -INFINITY, -INFINITY)}; }
            for (int i = 0; i < nk; ++i) {
                nd.box[i] = to_box(n2[kids[i]].box);
                if (n2[kids[i]].count > 0) { leaves.push_back(Leaf{n2[kids[i]].first, n2[kids[i]].count}); nd.child[i] = ~((int)leaves.size() - 1); }
                else { nodes.push_back(Node{}); nd.child[i] = (int)nodes.size(); st.push_back(Work{kids[i], (int)nodes.size() - 1}); }
            }
            nodes[w.n4] = nd;
        }
    }
    // Generic ordered walk: visit(leaf) is called for leaves whose (conservatively widened) box the ray enters before `closest()`.
    template <class Visit, class Closest, class Done>
    void walk(const Ray& ray, Visit visit, Closest closest, Done done) const {
        if (nodes.empty()) return;
        struct E { int node; float t; };
        E stack[128]; int sp = 0;
        stack[sp++] = E{1, ray.tmin};
        const float k = 9.5367431640625e-07f;
        while (sp && !done()) {
            const E top = stack[--sp];
            if (top.t > closest() * 1.0000152587890625f) continue;                                  // mapping_cpu.art:332 (conservative form)
            if (top.node < 0) { visit(leaves[~top.node]); continue; }
            const Node& n = nodes[top.node - 1];
            for (int i = 0; i < 4; ++i) {
                if (n.child[i] == 0) break;
                // slab test widened by its rounding error (cull_box), entry distance kept for the ordering
                float entry = ray.tmin, exit = closest() * 1.0000152587890625f;
                auto axis = [&](float inv_dir, float inv_org, float lo, float hi) {
                    const float a = fabsf(inv_org);
                    if (a == INFINITY) return;
                    const float t0 = fmaf_(inv_dir, lo, inv_org), t1 = fmaf_(inv_dir, hi, inv_org);
                    const float tn = fmin_sel(t0, t1) - a * k, tf = fmax_sel(t0, t1) + a * k;
                    if (tn > entry) entry = tn;
                    if (tf < exit) exit = tf;
                };
                axis(ray.inv_dir.x, ray.inv_org.x, n.box[i].min.x, n.box[i].max.x);
                axis(ray.inv_dir.y, ray.inv_org.y, n.box[i].min.y, n.box[i].max.y);
                axis(ray.inv_dir.z, ray.inv_org.z, n.box[i].min.z, n.box[i].max.z);
                if (exit < entry || sp >= 127) continue;
                // nearest on top: push, or slip under the current top (stack.art push_after; mapping_cpu.art:371-376)
                if (sp == 0 || stack[sp - 1].t > entry) stack[sp++] = E{n.child[i], entry};
                else { stack[sp] = stack[sp - 1]; stack[sp - 1] = E{n.child[i], entry}; ++sp; }
            }
        }
    }
};

struct MeshView {
    int num_face, num_verts, num_norms, num_tex;
    const float* verts;   // xyz + pad
    const float* norms;   // xyz + pad
    const int32_t* inds;  // i0 i1 i2 0
    const float* tex;     // uv
    Vec3 vertex(int i) const { return v3(verts[4 * i], verts[4 * i + 1], verts[4 * i + 2]); }
    Vec3 normal(int i) const { return v3(norms[4 * i], norms[4 * i + 1], norms[4 * i + 2]); }
    Vec2 texc(int i) const { return Vec2{tex[2 * i], tex[2 * i + 1]}; }
};

struct Shape {
    int type;
    MeshView mesh;            // trimesh
    std::vector<Tri> tris;    // per primitive, runtime/bvh/TriBVHAdapter.h:40-61 (p0, e1=p2-p0, e2=p0-p1, stable n)
    Bvh2 bvh;
    Bvh4 bvh4;                // timing arm
    Vec3 sph_origin; float sph_radius;
};

struct Entity { int id, shape_id, mat_id; Mat3x4 local_mat, global_mat; Mat3x3 normal_mat; };

// src/runtime/bvh/TriBVHAdapter.h:40-50
inline Vec3 stable_normal(Vec3 a, Vec3 b, Vec3 c) {
    const float ab_x = a.z * b.y, ab_y = a.x * b.z, ab_z = a.y * b.x;
    const float bc_x = b.z * c.y, bc_y = b.x * c.z, bc_z = b.y * c.x;
    const Vec3 cab = v3(a.y * b.z - ab_x, a.z * b.x - ab_y, a.x * b.y - ab_z);
    const Vec3 cbc = v3(b.y * c.z - bc_x, b.z * c.x - bc_y, b.x * c.y - bc_z);
    return v3(fabsf(ab_x) < fabsf(bc_x) ? cab.x : cbc.x, fabsf(ab_y) < fabsf(bc_y) ? cab.y : cbc.y, fabsf(ab_z) < fabsf(bc_z) ? cab.z : cbc.z);
}

struct Scene {
    std::vector<Entity> entities;
    std::vector<EntityLeaf> leaves;
    std::vector<Shape> shapes;
    std::vector<int> entity_per_material;
    std::vector<MaterialDesc> materials;
    std::vector<LightDesc> inf_lights, fin_lights;
    std::vector<uint8_t> shape_blob;
    CameraDesc camera; TechniqueDesc technique;
    std::vector<float> selector_data;
    std::vector<TextureDesc> textures;
    std::vector<ImageDesc> images;                  // pixels point into image_data
    std::vector<std::vector<uint8_t>> image_data;
    std::vector<float> aux_data;
    BBox bbox;
    Bvh2 top;   // over leaves
    Bvh4 top4;  // timing arm
    int num_materials;

    explicit Scene(const SceneDesc& d) {
        shape_blob.assign(d.shape_data, d.shape_data + d.shape_data_bytes);
        // driver/entity.art:12-29
        for (int i = 0; i < d.n_entities; ++i) {
            const float* r = d.entities + 36 * i;
            Entity e;
            e.id = i;
            auto c = [&](int o) { return v3(r[o], r[o + 1], r[o + 2]); };
            e.local_mat = Mat3x4{c(0), c(3), c(6), c(9)};
            e.global_mat = Mat3x4{c(12), c(15), c(18), c(21)};
            e.normal_mat = Mat3x3{c(24), c(27), c(30)};
            std::memcpy(&e.shape_id, r + 33, 4);
            std::memcpy(&e.mat_id, r + 34, 4);
            entities.push_back(e);
        }
        leaves.assign(d.leaves, d.leaves + d.n_leaves);
        for (int s = 0; s < d.n_shapes; ++s) {
            Shape sh{};
            sh.type = (int)d.shape_lookups[s].type_id;
            const uint8_t* p = shape_blob.data() + d.shape_lookups[s].offset;
            if (sh.type == SHAPE_TRIMESH) {
                // shapes/trimesh.art:77-96
                const int32_t* h = (const int32_t*)p;
                MeshView& m = sh.mesh;
                m.num_face = h[0]; m.num_verts = h[1]; m.num_norms = h[2]; m.num_tex = h[3];
                const float* f = (const float*)p;
                m.verts = f + 12;
                m.norms = m.verts + 4 * m.num_verts;
                m.inds = (const int32_t*)(m.norms + 4 * m.num_norms);
                m.tex = (const float*)(m.inds + 4 * m.num_face);
                std::vector<BBox> boxes(m.num_face);
                sh.tris.resize(m.num_face);
                for (int t = 0; t < m.num_face; ++t) {
                    const Vec3 p0 = m.vertex(m.inds[4 * t]), p1 = m.vertex(m.inds[4 * t + 1]), p2 = m.vertex(m.inds[4 * t + 2]);
                    Tri tr;
                    tr.v0 = p0; tr.e1 = p2 - p0; tr.e2 = p0 - p1;
                    tr.n = stable_normal(tr.e1, tr.e2, p1 - p2);
                    sh.tris[t] = tr;
                    boxes[t].min = v3(std::min({p0.x, p1.x, p2.x}), std::min({p0.y, p1.y, p2.y}), std::min({p0.z, p1.z, p2.z}));
                    boxes[t].max = v3(std::max({p0.x, p1.x, p2.x}), std::max({p0.y, p1.y, p2.y}), std::max({p0.z, p1.z, p2.z}));
                }
                sh.bvh.build(boxes, 4);
                sh.bvh4.build(boxes);
            } else {
                const float* f = (const float*)p;  // shapes/sphere.art:74-81
                sh.sph_origin = v3(f[0], f[1], f[2]);
                sh.sph_radius = f[3];
            }
            shapes.push_back(std::move(sh));
        }
        // MeshView pointers refer to shape_blob which never reallocates after this point
        entity_per_material.assign(d.entity_per_material, d.entity_per_material + d.n_materials);
        materials.assign(d.materials, d.materials + d.n_materials);
        inf_lights.assign(d.infinite_lights, d.infinite_lights + d.n_infinite);
        fin_lights.assign(d.finite_lights, d.finite_lights + d.n_finite);
        camera = d.camera; technique = d.technique; num_materials = d.n_materials;
        if (d.selector_data && d.n_selector_data > 0) selector_data.assign(d.selector_data, d.selector_data + d.n_selector_data);
        if (d.textures && d.n_textures > 0) textures.assign(d.textures, d.textures + d.n_textures);
        if (d.aux_data && d.n_aux_data > 0) aux_data.assign(d.aux_data, d.aux_data + d.n_aux_data);
        for (int i = 0; i < d.n_images; ++i) {
            ImageDesc im = d.images[i];
            const size_t bpp = im.format == IMAGE_RGBA8 ? 4 : im.format == IMAGE_MONO8 ? 1 : 16;
            const uint8_t* px = (const uint8_t*)im.pixels;
            image_data.emplace_back(px, px + bpp * (size_t)im.width * (size_t)im.height);
            im.pixels = image_data.back().data();
            images.push_back(im);
        }
        bbox = BBox{v3(d.bbox_min[0], d.bbox_min[1], d.bbox_min[2]), v3(d.bbox_max[0], d.bbox_max[1], d.bbox_max[2])};
        std::vector<BBox> lb(leaves.size());
        for (size_t i = 0; i < leaves.size(); ++i) lb[i] = BBox{v3(leaves[i].min[0], leaves[i].min[1], leaves[i].min[2]), v3(leaves[i].max[0], leaves[i].max[1], leaves[i].max[2])};
        top.build(lb, 1);
        top4.build(lb);
    }
};

// ------------------------------------------------------------------------------------------ traversal
// Candidate ordering (see header): smaller distance wins; equal distance -> larger (entity, primitive).
inline bool better(float t, int ent, int prim, const Hit& h) {
    if (t < h.distance) return true;
    if (t > h.distance) return false;
    if (h.prim_id < 0) return true;  // "t <= tmax" accepts a hit exactly at the initial bound (intersection.art:96-98)
    return ent > h.ent_id || (ent == h.ent_id && prim > h.prim_id);
}

// One entity: traversal/mapping_cpu.art:282-412 (bottom level) + shapes/trimesh.art:124-144 / sphere.art:138-147
inline void intersect_entity(const Scene& sc, const EntityLeaf& leaf, const Ray& ray, bool any_hit, int use_bvh, Hit& hit, bool& done) {
    const int ent = leaf.entity_id & 0x7FFFFFFF;
    const Mat3x4 local{v3(leaf.local[0], leaf.local[1], leaf.local[2]), v3(leaf.local[3], leaf.local[4], leaf.local[5]),
                       v3(leaf.local[6], leaf.local[7], leaf.local[8]), v3(leaf.local[9], leaf.local[10], leaf.local[11])};
    const Ray lray = transform_ray(ray, local);   // traversal/mapping_cpu.art:484
    const Shape& sh = sc.shapes[leaf.shape_id];
    if (sh.type == SHAPE_SPHERE) {
        float t; Vec2 uv;
        if (intersect_sphere(sh.sph_origin, sh.sph_radius, lray, t, uv) && better(t, ent, 0, hit)) {
            hit = Hit{t, uv, 0, ent};
            if (any_hit) done = true;
        }
        return;
    }
    auto test_tri = [&](int prim) {
        float t, u, v;
        if (intersect_ray_tri_mt(false, lray, sh.tris[prim], t, u, v) && better(t, ent, prim, hit)) {
            hit = Hit{t, Vec2{u, v}, prim, ent};
            if (any_hit) done = true;
        }
    };
    if (!use_bvh) {
        for (int p = 0; p < sh.mesh.num_face && !done; ++p) test_tri(p);
        return;
    }
    if (use_bvh == 2) {
        sh.bvh4.walk(lray, [&](const Bvh4::Leaf& l) { for (int i = 0; i < l.count && !done; ++i) test_tri(sh.bvh4.order[l.first + i]); },
                     [&]() { return hit.distance; }, [&]() { return done; });
        return;
    }
    int stack[128]; int sp = 0;
    if (sh.bvh.nodes.empty()) return;
    stack[sp++] = 0;
    while (sp && !done) {
        const Bvh2::Node& n = sh.bvh.nodes[stack[--sp]];
        if (cull_box(lray, n.box, hit.distance)) continue;
        if (n.count) { for (int i = 0; i < n.count && !done; ++i) test_tri(sh.bvh.order[n.first + i]); }
        else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
}

// Top level: traversal/mapping_cpu.art:421-518
inline Hit traverse(const Scene& sc, const Ray& ray, bool any_hit, int use_bvh) {
    Hit hit = invalid_hit(ray.tmax);
    bool done = false;
    auto visit_leaf = [&](const EntityLeaf& leaf) {
        if (!check_ray_visibility(ray, leaf.flags)) return;                        // :479
        float en, ex;                                                               // :480 intersect_ray_box_single_section
        intersect_ray_box(ray, BBox{v3(leaf.min[0], leaf.min[1], leaf.min[2]), v3(leaf.max[0], leaf.max[1], leaf.max[2])}, ray.tmax, en, ex);
        if (!((en <= ex) & (ex >= 0))) return;
        // :481 culls on `tmin <= hit.distance`; with rounding that makes the result depend on the visiting order
        // (header, "ties"), so the running hit is only ever used through cull_box below
        intersect_entity(sc, leaf, ray, any_hit, use_bvh, hit, done);
    };
    if (!use_bvh) {
        for (size_t i = 0; i < sc.leaves.size() && !done; ++i) visit_leaf(sc.leaves[i]);
        return hit;
    }
    if (use_bvh == 2) {
        sc.top4.walk(ray, [&](const Bvh4::Leaf& l) { for (int i = 0; i < l.count && !done; ++i) visit_leaf(sc.leaves[sc.top4.order[l.first + i]]); },
                     [&]() { return hit.distance; }, [&]() { return done; });
        return hit;
    }
    if (sc.top.nodes.empty()) return hit;
    int stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp && !done) {
        const Bvh2::Node& n = sc.top.nodes[stack[--sp]];
        if (cull_box(ray, n.box, hit.distance)) continue;
        if (n.count) { for (int i = 0; i < n.count && !done; ++i) visit_leaf(sc.leaves[sc.top.order[n.first + i]]); }
        else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
    return hit;
}

// ------------------------------------------------------------------------------------------ shading structs
struct SurfaceElement { bool is_entering; Vec3 point, face_normal; float area, inv_area; Vec2 prim_coords, tex_coords; Mat3x3 local; };
enum PdfMeasure { PDF_SOLID, PDF_AREA, PDF_DELTA };
struct Pdf {  // driver/pdf.art:16-46
    float value; PdfMeasure m;
    float as_solid(float cos, float dist2) const { return m == PDF_AREA ? value * dist2 / cos : (m == PDF_DELTA ? 1.0f : value); }
};
struct DirectLightSample { Vec3 pos, dir; Color intensity; Pdf pdf; float cos, dist; };

// core/triangle.art:12-29,32-43
struct Triangle { Vec3 v0, v1, v2, n; float area; };
inline Triangle make_triangle(Vec3 v0, Vec3 v1, Vec3 v2) {
    const Vec3 e1 = v2 - v0, e2 = v0 - v1, e3 = v1 - v2;
    const float x12 = e1.z * e2.y, y12 = e1.x * e2.z, z12 = e1.y * e2.x;
    const float x23 = e2.z * e3.y, y23 = e2.x * e3.z, z23 = e2.y * e3.x;
    const Vec3 c12 = v3(e1.y * e2.z - x12, e1.z * e2.x - y12, e1.x * e2.y - z12);
    const Vec3 c23 = v3(e2.y * e3.z - x23, e2.z * e3.x - y23, e2.x * e3.y - z23);
    const Vec3 n = v3(fabsf(x12) < fabsf(x23) ? c12.x : c23.x, fabsf(y12) < fabsf(y23) ? c12.y : c23.y, fabsf(z12) < fabsf(z23) ? c12.z : c23.z);
    const float nn = len(n);
    return Triangle{v0, v1, v2, mulf(n, 1 / nn), nn / 2};
}

// shapes/trimesh.art:14-40
inline SurfaceElement trimesh_surface_element(const MeshView& m, const Entity& e, const Ray& ray, const Hit& hit) {
    const int i0 = m.inds[4 * hit.prim_id], i1 = m.inds[4 * hit.prim_id + 1], i2 = m.inds[4 * hit.prim_id + 2];
    const Triangle tri = make_triangle(transform_point(e.global_mat, m.vertex(i0)), transform_point(e.global_mat, m.vertex(i1)), transform_point(e.global_mat, m.vertex(i2)));
    const Vec3 face_normal = tri.n;
    const Vec3 normal = normalize(mat3x3_mul(e.normal_mat, lerp2(m.normal(i0), m.normal(i1), m.normal(i2), hit.prim_coords.x, hit.prim_coords.y)));
    const bool is_entering = dot(ray.dir, face_normal) <= 0;
    const Vec2 t0 = m.texc(i0), t1 = m.texc(i1), t2 = m.texc(i2);
    SurfaceElement s;
    s.is_entering = is_entering;
    s.point = ray.org + mulf(ray.dir, hit.distance);
    s.face_normal = is_entering ? face_normal : neg(face_normal);
    s.area = tri.area;
    s.inv_area = safe_div(1, tri.area);
    s.prim_coords = hit.prim_coords;
    s.tex_coords = Vec2{lerp2(t0.x, t1.x, t2.x, hit.prim_coords.x, hit.prim_coords.y), lerp2(t0.y, t1.y, t2.y, hit.prim_coords.x, hit.prim_coords.y)};
    s.local = make_orthonormal(is_entering ? normal : neg(normal));
    return s;
}
// shapes/trimesh.art:41-68
inline SurfaceElement trimesh_surface_element_for_point(const MeshView& m, const Entity& e, int prim_id, Vec2 pc) {
    const int i0 = m.inds[4 * prim_id], i1 = m.inds[4 * prim_id + 1], i2 = m.inds[4 * prim_id + 2];
    const Vec3 g0 = transform_point(e.global_mat, m.vertex(i0)), g1 = transform_point(e.global_mat, m.vertex(i1)), g2 = transform_point(e.global_mat, m.vertex(i2));
    const Triangle tri = make_triangle(g0, g1, g2);
    const Vec3 normal = normalize(mat3x3_mul(e.normal_mat, lerp2(m.normal(i0), m.normal(i1), m.normal(i2), pc.x, pc.y)));
    const Vec2 t0 = m.texc(i0), t1 = m.texc(i1), t2 = m.texc(i2);
    SurfaceElement s;
    s.is_entering = true;
    s.point = lerp2(g0, g1, g2, pc.x, pc.y);
    s.face_normal = tri.n;
    s.area = tri.area;
    s.inv_area = safe_div(1, tri.area);
    s.prim_coords = pc;
    s.tex_coords = Vec2{lerp2(t0.x, t1.x, t2.x, pc.x, pc.y), lerp2(t0.y, t1.y, t2.y, pc.x, pc.y)};
    s.local = make_orthonormal(normal);
    return s;
}
// shapes/sphere.art:52-76 (area is only consumed by sphere area lights, which are out of scope: left 0)
inline SurfaceElement sphere_surface_element(Vec3 origin, const Entity& e, const Ray& ray, const Hit& hit) {
    const Vec3 point = ray.org + mulf(ray.dir, hit.distance);
    const Vec3 dir = point - transform_point(e.global_mat, origin);
    const float l = len(dir);
    const Vec3 normal = mulf(dir, 1 / l);
    SurfaceElement s;
    s.is_entering = true;
    s.point = point; s.face_normal = normal; s.area = 0; s.inv_area = 0;
    s.prim_coords = hit.prim_coords; s.tex_coords = hit.prim_coords;
    s.local = make_orthonormal(normal);
    return s;
}

// core/sampling.art:13-21,62-69
struct DirSample { Vec3 dir; float pdf; };
inline DirSample sample_cosine_hemisphere(float u, float v) {
    const float c = safe_sqrt(v), s = safe_sqrt(1 - v);
    const float phi = 2 * flt_pi * u;
    float sn, cs; dm_sincosf(phi, &sn, &cs);
    return DirSample{v3(s * cs, s * sn, c), c / flt_pi};
}
// core/warp.art:63-91
inline Vec3 equal_area_square_to_sphere(float px, float py) {
    const float u = 2 * px - 1, v = 2 * py - 1;
    const float au = fabsf(u), av = fabsf(v);
    const float signedDistance = 1 - (au + av);
    const float d = fabsf(signedDistance);
    const float r = 1 - d;
    const float phi = (r == 0 ? 1.0f : (av - au) / r + 1) * flt_pi / 4;
    const float cosTheta = copysignf(1 - r * r, signedDistance);
    const float sinTheta = safe_sqrt(2 - r * r) * r;
    float sn, cs; dm_sincosf(phi, &sn, &cs);
    const float cosPhi = copysignf(cs, u), sinPhi = copysignf(sn, v);
    return v3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta);
}

// core/fresnel.art:7-27
struct FresnelTerm { float cos_t, factor; };
inline float fresnel_factor(float eta, float cos_i, float cos_t) {
    const float R_s = safe_div(eta * cos_i - cos_t, eta * cos_i + cos_t);
    const float R_p = safe_div(cos_i - eta * cos_t, cos_i + eta * cos_t);
    return clampf((R_s * R_s + R_p * R_p) * 0.5f, 0, 1);
}
inline bool fresnel(float eta, float cos_i, FresnelTerm& out) {
    const float eta2 = cos_i < 0 ? 1 / eta : eta;
    const float cos2_t = 1 - (1 - cos_i * cos_i) * eta2 * eta2;
    if (cos2_t <= 0.0f) return false;
    const float cos_t = sqrtf(cos2_t);
    out.cos_t = cos_i < 0 ? -cos_t : cos_t;
    out.factor = fresnel_factor(eta2, fabsf(cos_i), cos_t);
    return true;
}

// core/fresnel.art:29-36
inline float conductor_factor(float n, float k, float cos_i) {
    const float f = n * n + k * k;
    const float d1 = f * cos_i * cos_i;
    const float d2 = 2.0f * n * cos_i;
    const float R_s = safe_div(d1 - d2, d1 + d2);
    const float R_p = safe_div(f - d2 + cos_i * cos_i, f + d2 + cos_i * cos_i);
    return clampf((R_s * R_s + R_p * R_p) * 0.5f, 0, 1);
}

// ------------------------------------------------------------------------------------------ textures
// core/math.art:74,88-91
inline float fractf_(float x) { return x - floorf(x); }
inline float wrapf_(float v, float mn, float mx) { const float range = mx - mn; return range <= flt_eps ? mn : v - (range * floorf((v - mn) / range)); }
// core/matrix.art:237-240 on the two rows the generator inlines (LoaderUtils::inlineTransformAs2d)
inline Vec2 transform_point_affine(const float* m, Vec2 v) { return Vec2{dot(v3(m[0], m[1], m[2]), v3(v.x, v.y, 1)), dot(v3(m[3], m[4], m[5]), v3(v.x, v.y, 1))}; }
struct Color4 { float r, g, b, a; };
inline Color4 c4lerp(Color4 a, Color4 b, float t) { return Color4{(1 - t) * a.r + t * b.r, (1 - t) * a.g + t * b.g, (1 - t) * a.b + t * b.b, (1 - t) * a.a + t * b.a}; }   // core/color.art:17-21
inline Color4 c4mulf(Color4 a, float f) { return Color4{a.r * f, a.g * f, a.b * f, a.a * f}; }
inline Color4 c4add(Color4 a, Color4 b) { return Color4{a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a}; }
// texture/image.art:9-44
inline int border_index(int mode, int x, int w) {
    if (mode == BORDER_CLAMP) return x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
    if (mode == BORDER_MIRROR) { const int t = x < 0 ? -1 - x : x; const int i = t / w; const int k = t - i * w; return (i & 1) == 0 ? w - 1 - k : k; }
    const int t = x % w; return t < 0 ? t + w : t;
}
// driver/image.art:9-34
inline Color4 image_pixel(const ImageDesc& im, int x, int y) {
    const size_t i = (size_t)y * (size_t)im.width + (size_t)x;
    if (im.format == IMAGE_RGBA8) { const uint8_t* p = (const uint8_t*)im.pixels + 4 * i; return Color4{(float)p[0] / 255, (float)p[1] / 255, (float)p[2] / 255, (float)p[3] / 255}; }
    if (im.format == IMAGE_MONO8) { const float g = (float)((const uint8_t*)im.pixels)[i] / 255; return Color4{g, g, g, 1}; }
    const float* p = (const float*)im.pixels + 4 * i; return Color4{p[0], p[1], p[2], p[3]};
}
// texture/image.art:78-146
inline Color4 image_filter(const ImageDesc& im, int filter, int bu, int bv, Vec2 uv) {
    if (filter == FILTER_NEAREST) {
        const float u = uv.x * (float)im.width, v = uv.y * (float)im.height;
        return image_pixel(im, border_index(bu, (int)floorf(u), im.width), border_index(bv, (int)floorf(v), im.height));
    }
    const float u = uv.x * (float)im.width - 0.5f, v = uv.y * (float)im.height - 0.5f;
    const int ix = (int)floorf(u), iy = (int)floorf(v);
    const float fx = fractf_(u), fy = fractf_(v);
    if (filter == FILTER_BILINEAR) {
        const int x0 = border_index(bu, ix, im.width), y0 = border_index(bv, iy, im.height), x1 = border_index(bu, ix + 1, im.width), y1 = border_index(bv, iy + 1, im.height);
        return c4lerp(c4lerp(image_pixel(im, x0, y0), image_pixel(im, x1, y0), fx), c4lerp(image_pixel(im, x0, y1), image_pixel(im, x1, y1), fx), fy);
    }
    auto w0 = [](float a) { return (a * (a * (-a + 3) - 3) + 1) / 6; };
    auto w1 = [](float a) { return (a * a * (3 * a - 6) + 4) / 6; };
    auto w2 = [](float a) { return (a * (a * (-3 * a + 3) + 3) + 1) / 6; };
    auto w3 = [](float a) { return (a * a * a) / 6; };
    auto g0 = [&](float a) { return w0(a) + w1(a); };
    auto g1 = [&](float a) { return w2(a) + w3(a); };
    auto h0 = [&](float a) { return (w1(a) / g0(a)) - 1; };
    auto h1 = [&](float a) { return (w3(a) / g1(a)) + 1; };
    const float g0x = g0(fx), g0y = g0(fy), g1x = g1(fx), g1y = g1(fy);
    const int ix0 = (int)floorf((float)ix + h0(fx) + 0.5f), iy0 = (int)floorf((float)iy + h0(fy) + 0.5f);
    const int ix1 = (int)floorf((float)ix + h1(fx) + 0.5f), iy1 = (int)floorf((float)iy + h1(fy) + 0.5f);
    const int x0 = border_index(bu, ix0, im.width), y0 = border_index(bv, iy0, im.height), x1 = border_index(bu, ix1, im.width), y1 = border_index(bv, iy1, im.height);
    const Color4 p00 = c4mulf(image_pixel(im, x0, y0), g0x * g0y), p10 = c4mulf(image_pixel(im, x1, y0), g1x * g0y);
    const Color4 p01 = c4mulf(image_pixel(im, x0, y1), g0x * g1y), p11 = c4mulf(image_pixel(im, x1, y1), g1x * g1y);
    return c4add(c4add(p00, p10), c4add(p01, p11));
}
struct TextureSet { const TextureDesc* tex; const ImageDesc* img; };
// texture/checkerboard.art:4-13, texture/image.art:148-153; uv = ctx.uvw.xy
inline Color eval_texture(const TextureSet& ts, int id, Vec2 uv) {
    const TextureDesc& t = ts.tex[id];
    const Vec2 uv2 = transform_point_affine(t.transform, uv);
    if (t.type == TEX_CHECKERBOARD) {
        const float sx = uv2.x * t.p[0], sy = uv2.y * t.p[1];
        const bool px = ((int)wrapf_(sx, 0, 2) % 2) == 0, py = ((int)wrapf_(sy, 0, 2) % 2) == 0;
        return (px != py) ? col(t.p[2], t.p[3], t.p[4]) : col(t.p[5], t.p[6], t.p[7]);
    }
    const Color4 c = image_filter(ts.img[t.image], t.filter, t.border_u, t.border_v, uv2);
    return col(c.r, c.g, c.b);
}

// ------------------------------------------------------------------------------------------ microfacets (core/microfacet.art)
inline float absolute_cos(Vec3 a, Vec3 b) { return fabsf(dot(a, b)); }                                   // core/common.art:270
inline Vec3 reflect(Vec3 v, Vec3 n) { return mulf(n, 2 * dot(n, v)) - v; }                                // core/vector.art:124
inline Vec3 to_world(const Mat3x3& l, Vec3 v) { return (mulf(l.c0, v.x) + mulf(l.c1, v.y)) + mulf(l.c2, v.z); }   // core/shading.art:8
inline Vec3 to_local(const Mat3x3& l, Vec3 v) { return v3(dot(l.c0, v), dot(l.c1, v), dot(l.c2, v)); }           // core/shading.art:9 (= mat3x3_left_mul)
// :159-178
inline float g_1_smith(const Mat3x3& local, Vec3 w, float alpha_u, float alpha_v) {
    const float cosZ = dot(local.c2, w);
    if (fabsf(cosZ) <= flt_eps) return 0;
    const float cosX = dot(local.c0, w), cosY = dot(local.c1, w);
    const float kx = alpha_u * cosX, ky = alpha_v * cosY;
    const float a2 = kx * kx + ky * ky;
    if (a2 <= flt_eps) return 1;
    const float k2 = a2 / (cosZ * cosZ);
    const float denom = 1 + sqrtf(1 + k2);
    return 2 / denom;
}
// :192-202
inline float ndf_ggx(const Mat3x3& local, Vec3 m, float alpha_u, float alpha_v) {
    const float cosZ = dot(local.c2, m), cosX = dot(local.c0, m), cosY = dot(local.c1, m);
    const float kx = cosX / alpha_u, ky = cosY / alpha_v;
    const float k = kx * kx + ky * ky + cosZ * cosZ;
    return safe_div(1, flt_pi * alpha_u * alpha_v * k * k);
}
// :370-394 (Dupuy & Benyoub, spherical caps), as written: the UNSTRETCHED view vector is added to the cap sample
inline Vec3 sample_vndf_ggx(Rng& rnd, const Mat3x3& local, Vec3 vN, float alpha_u, float alpha_v) {
    const Vec3 vL = to_local(local, vN);
    const Vec3 sL = normalize(v3(alpha_u * vL.x, alpha_v * vL.y, vL.z));
    const float u0 = rnd.next_f32(); const float u1 = rnd.next_f32();
    const float phi = 2 * flt_pi * u0;
    const float z = (1 - u1) * (1 + sL.z) - sL.z;
    const float sinTheta = sqrtf(clampf(1 - z * z, 0, 1));
    float sn, cs; dm_sincosf(phi, &sn, &cs);
    const float x = sinTheta * cs, y = sinTheta * sn;
    const Vec3 h = v3(x, y, z) + vL;
    const Vec3 Nh = normalize(v3(h.x * alpha_u, h.y * alpha_v, h.z));
    return to_world(local, Nh);
}
// :396-399
inline float pdf_vndf_ggx(const Mat3x3& local, Vec3 w, Vec3 h, float alpha_u, float alpha_v) {
    const float cosZ = absolute_cos(local.c2, w);
    return safe_div(g_1_smith(local, w, alpha_u, alpha_v) * absolute_cos(w, h) * ndf_ggx(local, h, alpha_u, alpha_v), cosZ);
}
inline bool check_if_delta_distribution(float au, float av) { return au <= 1e-4f || av <= 1e-4f; }       // :297

// ------------------------------------------------------------------------------------------ normal / bump mapping (bsdf/map.art, core/sampling.art:118-165, core/matrix.art:261-284)
inline Vec3 ensure_valid_reflection(Vec3 Ng, Vec3 I, Vec3 N) {
    const Vec3 R = reflect(I, N);
    const float threshold = fminf(0.9f * dot(Ng, I), 0.01f);
    if (dot(Ng, R) >= threshold) return N;
    const float NdotNg = dot(N, Ng);
    const Vec3 X = normalize(N - mulf(Ng, NdotNg));
    const float Ix = dot(I, X), Iz = dot(I, Ng);
    const float Ix2 = Ix * Ix, Iz2 = Iz * Iz;
    const float a = Ix2 + Iz2;
    const float b = safe_sqrt(Ix2 * (a - threshold * threshold));
    const float c = Iz * threshold + a;
    const float fac = 0.5f / a;
    const float N1_z2 = fac * (b + c), N2_z2 = fac * (-b + c);
    const bool valid1 = (N1_z2 > 1e-5f) && (N1_z2 <= (1.0f + 1e-5f)), valid2 = (N2_z2 > 1e-5f) && (N2_z2 <= (1.0f + 1e-5f));
    Vec2 Nn;
    if (valid1 && valid2) {
        const Vec2 N1{safe_sqrt(1 - N1_z2), safe_sqrt(N1_z2)}, N2{safe_sqrt(1 - N2_z2), safe_sqrt(N2_z2)};
        const float R1 = 2 * (N1.x * Ix + N1.y * Iz) * N1.y - Iz, R2 = 2 * (N2.x * Ix + N2.y * Iz) * N2.y - Iz;
        const bool valid3 = R1 >= 1e-5f, valid4 = R2 >= 1e-5f;
        if (valid3 && valid4) Nn = R1 < R2 ? N1 : N2; else Nn = R1 > R2 ? N1 : N2;
    } else if (valid1 || valid2) {
        const float Nz2 = valid1 ? N1_z2 : N2_z2;
        Nn = Vec2{safe_sqrt(1 - Nz2), safe_sqrt(Nz2)};
    } else Nn = Vec2{0, 1};
    return mulf(X, Nn.x) + mulf(Ng, Nn.y);
}
inline Mat3x3 mat3x3_align_vectors(Vec3 a, Vec3 b) {
    const Vec3 axis = cross(b, a);
    const float cosA = dot(a, b);
    if (cosA <= -1) return Mat3x3{v3(-1, 0, 0), v3(0, -1, 0), v3(0, 0, -1)};
    const float k = 1 / (1 + cosA);
    return Mat3x3{v3((axis.x * axis.x * k) + cosA, (axis.y * axis.x * k) - axis.z, (axis.z * axis.x * k) + axis.y),
                  v3((axis.x * axis.y * k) + axis.z, (axis.y * axis.y * k) + cosA, (axis.z * axis.y * k) - axis.x),
                  v3((axis.x * axis.z * k) - axis.y, (axis.y * axis.z * k) + axis.x, (axis.z * axis.z * k) + cosA)};
}
inline Mat3x3 mat3x3_matmul(const Mat3x3& a, const Mat3x3& b) { return Mat3x3{mat3x3_mul(a, b.c0), mat3x3_mul(a, b.c1), mat3x3_mul(a, b.c2)}; }   // core/matrix.art:124-127
// make_normal_set, bsdf/map.art:39-45: the frame the inner BSDF is built on
inline Mat3x3 normal_set_frame(const SurfaceElement& surf, Vec3 ray_dir, Vec3 normal) {
    const Vec3 n = ensure_valid_reflection(surf.face_normal, neg(ray_dir), normalize(normal));
    return mat3x3_matmul(mat3x3_align_vectors(surf.local.c2, n), surf.local);
}

// ------------------------------------------------------------------------------------------ BSDFs
struct BsdfSample { Vec3 in_dir; float pdf; Color color; float eta; bool is_delta; };
struct Bsdf {
    int type; Mat3x3 local; bool is_entering; Color kd; float n1, n2; Color ks, kt; Color c_eta, c_k; bool mirror;
    bool rough; float alpha_u, alpha_v;   // CONDUCTOR with make_vndf_ggx_distribution (bsdf/conductor.art:45-141)
    bool is_all_delta() const { return type == BSDF_DIELECTRIC || (type == BSDF_CONDUCTOR && !rough); }
    Color fresnel_term(float c) const { return col(conductor_factor(c_eta.r, c_k.r, c), conductor_factor(c_eta.g, c_k.g, c), conductor_factor(c_eta.b, c_k.b, c)); }
    // bsdf/diffuse.art:2-12 ; bsdf/dielectric.art:15-37 ; bsdf/conductor.art:78-91
    Color eval(Vec3 in_dir, Vec3 out_dir) const {
        if (type == BSDF_DIFFUSE) return cmulf(kd, positive_cos(in_dir, local.c2) * flt_inv_pi);
        if (type == BSDF_CONDUCTOR && rough) {
            const Vec3 N = local.c2;
            const float cos_o = absolute_cos(out_dir, N), cos_i = absolute_cos(in_dir, N);
            if (cos_o <= flt_eps || cos_i <= flt_eps) return col(0, 0, 0);
            const Vec3 H = normalize(in_dir + out_dir);
            const float D = ndf_ggx(local, H, alpha_u, alpha_v);
            const float G = g_1_smith(local, in_dir, alpha_u, alpha_v) * g_1_smith(local, out_dir, alpha_u, alpha_v);
            const Color F = fresnel_term(absolute_cos(out_dir, H));
            const Color IF = col(1 - F.r, 1 - F.g, 1 - F.b);
            return cmulf(cadd(cmul(col(0, 0, 0), IF), cmul(ks, F)), D * G / (4 * cos_o));   // kd = black (make_rough_conductor_bsdf)
        }
        return col(0, 0, 0);
    }
    float pdf(Vec3 in_dir, Vec3 out_dir) const {
        if (type == BSDF_DIFFUSE) return positive_cos(in_dir, local.c2) / flt_pi;
        if (type == BSDF_CONDUCTOR && rough) {   // conductor.art:95-100
            const Vec3 H = normalize(in_dir + out_dir);
            const float cos_h_o = absolute_cos(out_dir, H);
            const float jacob = safe_div(1, 4 * cos_h_o);
            return pdf_vndf_ggx(local, out_dir, H, alpha_u, alpha_v) * jacob;
        }
        return 0.0f;
    }
    bool sample(Rng& rnd, Vec3 out_dir, bool adjoint, BsdfSample& s) const {
        if (type == BSDF_DIFFUSE) {
            const float u = rnd.next_f32(); const float v = rnd.next_f32();
            const DirSample ds = sample_cosine_hemisphere(u, v);
            s = BsdfSample{mat3x3_mul(local, ds.dir), ds.pdf, kd, 1, false};
            return true;
        }
        if (type == BSDF_CONDUCTOR && rough) {   // conductor.art:101-122
            const Vec3 N = local.c2;
            const float cos_o = absolute_cos(out_dir, N);
            if (cos_o <= flt_eps) return false;
            const Vec3 m = sample_vndf_ggx(rnd, local, out_dir, alpha_u, alpha_v);
            const float m_pdf = pdf_vndf_ggx(local, out_dir, m, alpha_u, alpha_v);
            if (len2(m) <= flt_eps) return false;
            const Vec3 oH = normalize(m);
            const Vec3 H = std::signbit(dot(oH, out_dir)) ? neg(oH) : oH;
            const Vec3 in_dir = reflect(out_dir, H);
            const float cos_i = absolute_cos(in_dir, N);
            if (cos_i <= flt_eps) return false;
            const float cos_h_o = absolute_cos(out_dir, H);
            const float jacob = 1 / (4 * cos_h_o);
            const float pdf = m_pdf * jacob;
            s = BsdfSample{in_dir, pdf, cmulf(eval(in_dir, out_dir), safe_div(1, pdf)), 1, false};
            return true;
        }
        if (type == BSDF_CONDUCTOR) {
            // bsdf/conductor.art:2-10 (make_mirror_bsdf) and :14-27 (make_pure_conductor_bsdf); which one is the generator's
            // partial-evaluation decision (conductor.art:131-141: eta, k known constants ~ (0, 1))
            const Vec3 nn = local.c2;
            const Vec3 r = mulf(nn, 2 * dot(nn, out_dir)) - out_dir;   // core/vector.art:124
            if (mirror) { s = BsdfSample{r, 1, ks, 1, true}; return true; }
            const float cos_i = dot(out_dir, nn);
            s = BsdfSample{r, 1, cmul(ks, fresnel_term(cos_i)), 1, true};
            return true;
        }
        const float k = is_entering ? n1 / n2 : n2 / n1;
        const Vec3 n = local.c2;
        const float cos_o = dot(out_dir, n);
        FresnelTerm ft{0, 1};
        if (!fresnel(k, cos_o, ft)) ft = FresnelTerm{0, 1};
        if (rnd.next_f32() > ft.factor) {
            // core/vector.art:127 vec3_refract
            const Vec3 t = mulf(n, k * cos_o - ft.cos_t) - mulf(out_dir, k);
            const float adj = adjoint ? k * k : 1.0f;
            s = BsdfSample{t, 1, cmulf(kt, adj), k, true};
        } else {
            // core/vector.art:124 vec3_reflect
            s = BsdfSample{mulf(n, 2 * dot(n, out_dir)) - out_dir, 1, ks, 1, true};
        }
        return true;
    }
    // Bsdf.albedo, for the Albedo AOV (technique/internal/infobuffer.art): diffuse.art:10, dielectric.art:35, conductor.art:9,28-38,50-56
    Color albedo(Vec3 out_dir) const {
        if (type == BSDF_DIFFUSE) return kd;
        if (type == BSDF_DIELECTRIC) return col(lerp(ks.r, kt.r, 0.5f), lerp(ks.g, kt.g, 0.5f), lerp(ks.b, kt.b, 0.5f));
        if (rough) { const Color F = fresnel_term(absolute_cos(out_dir, local.c2)); return cadd(cmul(col(0, 0, 0), col(1 - F.r, 1 - F.g, 1 - F.b)), cmul(ks, F)); }
        if (mirror) return ks;
        return cmul(ks, fresnel_term(dot(out_dir, local.c2)));
    }
};

// ------------------------------------------------------------------------------------------ lights
struct LightCtx { const Scene* sc; };

// light/area.art:124-258 plane emitter
struct SQ { Vec3 o, n; float x0, y0, z0, x1, y1, b0, b1, k, s; };
struct PlaneEmitter {
    Vec3 origin, x_axis, y_axis, normal; float area, inv_area, width, height; Vec3 ex, ey; Vec2 t0, t1, t2, t3;
    explicit PlaneEmitter(const LightDesc& l) {
        const float* p = l.p;
        origin = v3(p[0], p[1], p[2]); x_axis = v3(p[3], p[4], p[5]); y_axis = v3(p[6], p[7], p[8]); normal = v3(p[9], p[10], p[11]);
        area = p[12];
        t0 = Vec2{p[13], p[14]}; t1 = Vec2{p[15], p[16]}; t2 = Vec2{p[17], p[18]}; t3 = Vec2{p[19], p[20]};
        inv_area = safe_div(1, area);
        width = len(x_axis); height = len(y_axis);
        ex = mulf(x_axis, 1 / width); ey = mulf(y_axis, 1 / height);
    }
    SQ compute_sq(Vec3 from_point) const {
        const Vec3 dir = origin - from_point;
        const float x0 = dot(dir, ex), y0 = dot(dir, ey), z0_ = dot(dir, normal);
        const float x1 = x0 + width, y1 = y0 + height;
        const bool pos = !std::signbit(z0_);
        const float z0 = pos ? -z0_ : z0_;
        const Vec3 n = pos ? neg(normal) : normal;
        // diff = (x0,y1,x1,y0) - (x1,y0,x0,y1); nz_ = (y0,x1,y1,x0) * diff
        const float df[4] = {x0 - x1, y1 - y0, x1 - x0, y0 - y1};
        const float a[4] = {y0, x1, y1, x0};
        float nz[4];
        for (int i = 0; i < 4; ++i) {
            const float nz_ = a[i] * df[i];
            nz[i] = nz_ / sqrtf((df[i] * df[i]) * (z0 * z0) + nz_ * nz_);
        }
        auto safe_acos = [](float x) { return dm_acosf(clampf(x, -1, 1)); };
        const float g0 = safe_acos(-nz[0] * nz[1]), g1 = safe_acos(-nz[1] * nz[2]), g2 = safe_acos(-nz[2] * nz[3]), g3 = safe_acos(-nz[3] * nz[0]);
        SQ q;
        q.o = from_point; q.n = n; q.x0 = x0; q.y0 = y0; q.z0 = z0; q.x1 = x1; q.y1 = y1;
        q.b0 = nz[0]; q.b1 = nz[2];
        q.k = 2 * flt_pi - g2 - g3;
        q.s = g0 + g1 - q.k;
        return q;
    }
    void sample_direct(Vec2 uv, Vec3 from_point, SurfaceElement& surf, Pdf& pdf, float& weight) const {
        const SQ sq = compute_sq(from_point);
        const float au = fmaf_(uv.x, sq.s, sq.k);
        float sn, cs; dm_sincosf(au, &sn, &cs);
        const float fu = fmaf_(cs, sq.b0, -sq.b1) / sn;
        const float cu = clampf(copysignf(1.0f, fu) / sqrtf(sum_of_prod(fu, fu, sq.b0, sq.b0)), -1, 1);
        const float xu = clampf(-(cu * sq.z0) / sqrtf(fmaf_(-cu, cu, 1.0f)), sq.x0, sq.x1);
        const float d = sqrtf(sum_of_prod(xu, xu, sq.z0, sq.z0));
        const float h0 = sq.y0 / sqrtf(sum_of_prod(d, d, sq.y0, sq.y0));
        const float h1 = sq.y1 / sqrtf(sum_of_prod(d, d, sq.y1, sq.y1));
        const float hv = fmaf_(uv.y, h1 - h0, h0);
        const float hv2 = hv * hv;
        const float yv = (hv2 < 1 - 1e-6f) ? (hv * d) / sqrtf(1 - hv2) : sq.y1;
        const Vec3 p = sq.o + (mulf(ex, xu) + (mulf(ey, yv) + mulf(sq.n, sq.z0)));
        const float pdf_s = safe_div(1, sq.s);
        const float tx = dot(p - origin, ex) / width, ty = dot(p - origin, ey) / height;
        const Vec2 c0{lerp(t0.x, t1.x, tx), lerp(t0.y, t1.y, tx)}, c1{lerp(t2.x, t3.x, tx), lerp(t2.y, t3.y, tx)};
        surf.is_entering = true; surf.point = p; surf.face_normal = normal; surf.area = area; surf.inv_area = inv_area;
        surf.prim_coords = Vec2{tx, ty};
        surf.tex_coords = Vec2{lerp(c0.x, c1.x, ty), lerp(c0.y, c1.y, ty)};
        surf.local = make_orthonormal(normal);
        pdf = Pdf{pdf_s, PDF_SOLID};
        weight = sq.s;
    }
    Pdf pdf_direct(Vec3 from_point) const { const SQ sq = compute_sq(from_point); return Pdf{safe_div(1, sq.s), PDF_SOLID}; }
};

// light/area.art:62-107 shape emitter over a triangle mesh entity
inline void shape_emitter_sample(const Scene& sc, int entity_id, Vec2 uv, SurfaceElement& surf, float& pdfv, float& weight) {
    const Entity& e = sc.entities[entity_id];
    const Shape& sh = sc.shapes[e.shape_id];
    const int count = sh.mesh.num_face;
    const float ux = uv.x * (float)count;
    const int f = std::min((int)ux, count - 1);
    float u = ux - (float)f, v = uv.y;
    if (u + v > 1) { u = 1 - u; v = 1 - v; }  // core/sampling.art:34-36
    surf = trimesh_surface_element_for_point(sh.mesh, e, f, Vec2{u, v});
    pdfv = surf.inv_area / (float)count;
    weight = surf.area * (float)count;
}

// light/area.art:260-316 sphere emitter; p = radiance rgb, sphere origin (local) xyz, radius, area. The area is
// compute_ellipsoid_area (shapes/sphere.art:21-27), three pow() calls on per-light constants: evaluated once by the host
// (ignis_b200/scene.py, csrc/host/script_recognizer.cpp) and passed in, so that no transcendental sits on the parity path.
// shapes/sphere.art:29-45
inline SurfaceElement sphere_surface_for_normal(const Entity& e, Vec3 origin, float radius, float area, Vec3 normal) {
    const Vec3 point = origin + mulf(normal, radius);
    const Vec2 uv = sphere_map_uv(normal);
    const Vec3 gn = normalize(mat3x3_mul(e.normal_mat, normal));
    SurfaceElement s;
    s.is_entering = true; s.point = transform_point(e.global_mat, point); s.face_normal = gn; s.area = area; s.inv_area = safe_div(1, area);
    s.prim_coords = uv; s.tex_coords = uv; s.local = make_orthonormal(gn);
    return s;
}
inline void sphere_emitter_sample(const Scene& sc, const LightDesc& l, Vec2 uv, Vec3 from_point, SurfaceElement& surf, Pdf& pdf, float& weight) {
    const Entity& e = sc.entities[l.entity_id];
    const Vec3 origin = v3(l.p[3], l.p[4], l.p[5]); const float radius = l.p[6], area = l.p[7];
    const Vec3 glb_org = transform_point(e.global_mat, origin);
    surf = sphere_surface_for_normal(e, origin, radius, area, equal_area_square_to_sphere(uv.x, uv.y));
    const Vec3 p = surf.point;
    const Vec3 os = from_point - glb_org, ps = from_point - p;
    if (!(len2(ps) <= len2(os))) {   // the sampled point is on the far side: mirror it through the centre (:285-292)
        const Vec3 po = glb_org - p;
        const Vec3 np = p + mulf(po, 2);
        const Vec3 norm = normalize(np - glb_org);
        // pointmapper.art:33 to_local_normal = (normal_mat^T n) / |diag(normal_mat)|^2
        const Mat3x3& m = e.normal_mat;
        const Vec3 ln = mulf(v3(dot(m.c0, norm), dot(m.c1, norm), dot(m.c2, norm)), 1 / len2(v3(m.c0.x, m.c1.y, m.c2.z)));
        surf = sphere_surface_for_normal(e, origin, radius, area, ln);
    }
    pdf = Pdf{safe_div(1, area), PDF_AREA};
    weight = area;
}

struct LightRef { const LightDesc* d; bool infinite; int id; };

inline bool light_delta(const LightDesc& l) { return l.type == LIGHT_POINT || l.type == LIGHT_SPOT || l.type == LIGHT_DIRECTIONAL; }
inline bool light_infinite(const LightDesc& l) { return l.type == LIGHT_ENV_CONST || l.type == LIGHT_SUN || l.type == LIGHT_DIRECTIONAL || l.type == LIGHT_ENV_TEXTURED || l.type == LIGHT_ENV_TEX; }

// ---- 1-D / 2-D cdfs over a buffer without the leading 0 (core/cdf.art:34-75,105-155, core/interval.art:7-23)
struct Cdf1D {
    const float* data; int func_size;
    float get(int i) const { return i == 0 ? 0.0f : data[i - 1]; }
    float pdf_discrete(int x) const { return get(x + 1) - get(x); }
    int sample_discrete(float u, float& pdf) const {
        const int size = func_size + 1;
        int first = 0, len = size;
        while (len > 0) {
            const int half = len / 2, middle = first + half;
            if (get(middle) <= u) { first = middle + 1; len -= half + 1; } else len = half;
        }
        const int off = std::min(std::min(std::max(first - 1, 0), size - 1), func_size - 1);
        pdf = pdf_discrete(off);
        return off;
    }
    // returns the offset; pos in [0, 1], pdf with respect to pos
    int sample_continuous(float u, float& pos, float& pdf) const {
        float dpdf;
        const int off = sample_discrete(u, dpdf);
        const float rem = safe_div(u - get(off), dpdf);
        pos = clampf(((float)off + rem) / (float)func_size, 0, 1);
        pdf = dpdf * (float)func_size;
        return off;
    }
    int pdf_continuous(float x, float& pdf) const {
        const int off = std::min(std::max((int)(x * (float)func_size), 0), func_size - 1);
        pdf = pdf_discrete(off) * (float)func_size;
        return off;
    }
};
struct Cdf2D {   // make_cdf_2d_from_buffer: the marginal (over y) first, then size_y conditionals of size_x
    const float* data; int size_x, size_y;
    Cdf1D marginal() const { return Cdf1D{data, size_y}; }
    Cdf1D conditional(int i) const { return Cdf1D{data + size_y + (size_t)i * size_x, size_x}; }
    void sample_continuous(float ux, float uy, Vec2& pos, float& pdf) const {
        float p1, pdf1, p2, pdf2;
        const int off1 = marginal().sample_continuous(uy, p1, pdf1);
        conditional(off1).sample_continuous(ux, p2, pdf2);
        pos = Vec2{p2, p1}; pdf = pdf1 * pdf2;
    }
    float pdf_continuous(Vec2 pos) const {
        float pdf1, pdf2;
        const int off1 = marginal().pdf_continuous(pos.y, pdf1);
        conditional(off1).pdf_continuous(pos.x, pdf2);
        return pdf1 * pdf2;
    }
};
inline int32_t fbits(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline Vec3 switch_env_up(Vec3 v) { return v3(v.x, v.z, v.y); }                     // light/env.art:13
inline Vec2 map_env_uv(Vec3 dir) {                                                 // light/env.art:16-21, core/warp.art:44-48
    const float theta = dm_acosf(dir.z);
    float phi = dm_atan2f(dir.y, dir.x);
    if (phi < 0) phi = phi + 2 * flt_pi;
    const float v = theta / flt_pi, u = phi / (2 * flt_pi);
    return Vec2{fractf_(u + 0.25f), 1 - v};
}
inline Mat3x3 env_transform(const LightDesc& l) { return Mat3x3{v3(l.p[3], l.p[4], l.p[5]), v3(l.p[6], l.p[7], l.p[8]), v3(l.p[9], l.p[10], l.p[11])}; }
// make_environment_light_textured (light/env.art:112-160): direction pdf of `dir` and emitted radiance towards -dir
inline float env_textured_pdf(const LightDesc& l, const float* aux, Vec3 dir) {
    const Vec3 ldir = switch_env_up(mat3x3_mul(env_transform(l), dir));
    const float sinTheta = safe_sqrt(1 - ldir.z * ldir.z);                         // core/shading.art:16-17
    const Cdf2D cdf{aux + fbits(l.p[13]), fbits(l.p[14]), fbits(l.p[15])};
    return safe_div(cdf.pdf_continuous(map_env_uv(ldir)), sinTheta * flt_pi * flt_pi * 2);
}
inline Color env_textured_emission(const LightDesc& l, const TextureSet& ts, Vec3 dir) {
    const Vec3 ldir = switch_env_up(mat3x3_mul(env_transform(l), dir));
    return cmul(col(l.p[0], l.p[1], l.p[2]), eval_texture(ts, fbits(l.p[12]), map_env_uv(ldir)));
}
// make_environment_light over a texture (light/env.art:161-167): func(dir) = scale * tex(map_env_uv(switch_env_up(dir))), applied to transform * dir
inline Color env_tex_radiance(const LightDesc& l, const TextureSet& ts, Vec3 dir) {
    return cmul(col(l.p[0], l.p[1], l.p[2]), eval_texture(ts, fbits(l.p[12]), map_env_uv(switch_env_up(mat3x3_mul(env_transform(l), dir)))));
}
// core/warp.art:2-22
inline void square_to_concentric_disk(float px, float py, float& x, float& y) {
    const float a = 2 * px - 1, b = 2 * py - 1;
    if (a == 0 && b == 0) { x = 0; y = 0; return; }
    float sn, cs;
    if (a * a > b * b) { const float phi = (flt_pi / 4) * safe_div(b, a); dm_sincosf(phi, &sn, &cs); x = cs * a; y = sn * a; }
    else { const float phi = (flt_pi / 2) - (flt_pi / 4) * safe_div(a, b); dm_sincosf(phi, &sn, &cs); x = cs * b; y = sn * b; }
}
inline float uniform_cone_pdf(float cos_angle) { return safe_div(1, 2 * flt_pi * (1 - cos_angle)); }   // core/sampling.art:106
// core/sampling.art:109-116
inline DirSample sample_uniform_cone(float u, float v, float cos_angle) {
    const float c1 = 1 - cos_angle;
    float px, py; square_to_concentric_disk(u, v, px, py);
    const float n2 = px * px + py * py;
    const float z = cos_angle + c1 * (1 - n2);
    const float f = safe_sqrt(c1 * (2 - c1 * n2));
    return DirSample{v3(px * f, py * f, z), uniform_cone_pdf(cos_angle)};
}
inline bool sun_hit(const LightDesc& l, Vec3 dir) { return dot(v3(l.p[0], l.p[1], l.p[2]), dir) >= l.p[3]; }   // light/sun.art:18

inline DirectLightSample light_sample_direct(const Scene& sc, const LightDesc& l, Rng& rnd, const SurfaceElement& from) {
    switch (l.type) {
    case LIGHT_ENV_CONST: {  // light/env.art:84-88
        const float scene_radius = len(sc.bbox.max - sc.bbox.min) / 2 * 1.01f;  // core/bbox.art:24
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        const Vec3 dir = equal_area_square_to_sphere(u, v);
        const float pdf = 1 / (4 * flt_pi);
        const Color intensity = cmulf(col(l.p[0], l.p[1], l.p[2]), 1 / pdf);
        return DirectLightSample{from.point + mulf(dir, scene_radius), dir, intensity, Pdf{pdf, PDF_SOLID}, 1.0f, scene_radius};
    }
    case LIGHT_ENV_TEX: {  // light/env.art:84-88 with func = scale * tex
        const float scene_radius = len(sc.bbox.max - sc.bbox.min) / 2 * 1.01f;
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        const Vec3 dir = equal_area_square_to_sphere(u, v);
        const float pdf = 1 / (4 * flt_pi);
        const Color intensity = cmulf(env_tex_radiance(l, TextureSet{sc.textures.data(), sc.images.data()}, dir), 1 / pdf);
        return DirectLightSample{from.point + mulf(dir, scene_radius), dir, intensity, Pdf{pdf, PDF_SOLID}, 1.0f, scene_radius};
    }
    case LIGHT_ENV_TEXTURED: {  // light/env.art:115-126,136-139 (the sampled intensity is NOT multiplied by `scale`, as in the reference)
        const float scene_radius = len(sc.bbox.max - sc.bbox.min) / 2 * 1.01f;
        const Cdf2D cdf{sc.aux_data.data() + fbits(l.p[13]), fbits(l.p[14]), fbits(l.p[15])};
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        Vec2 pos; float spdf;
        cdf.sample_continuous(u, v, pos, spdf);
        const Color intensity = eval_texture(TextureSet{sc.textures.data(), sc.images.data()}, fbits(l.p[12]), pos);
        const float theta = (1 - pos.y) * flt_pi, phi = (pos.x - 0.25f) * 2 * flt_pi;
        float st, ct, sp, cp; dm_sincosf(theta, &st, &ct); dm_sincosf(phi, &sp, &cp);   // dir_from_spherical, core/warp.art:50-57
        const Vec3 d = v3(st * cp, st * sp, ct);
        const float sinTheta = safe_sqrt(1 - d.z * d.z);
        const float pdf_dir = safe_div(spdf, sinTheta * flt_pi * flt_pi * 2);
        const Vec3 dir = to_local(env_transform(l), switch_env_up(d));                  // mat3x3_left_mul(transform, .)
        return DirectLightSample{from.point + mulf(dir, scene_radius), dir, cmulf(intensity, 1 / pdf_dir), Pdf{pdf_dir, PDF_SOLID}, 1.0f, scene_radius};
    }
    case LIGHT_SUN: {  // light/sun.art:22-26
        const float cos_angle = l.p[3];
        const Mat3x3 frame = make_orthonormal(neg(v3(l.p[0], l.p[1], l.p[2])));
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        const DirSample smp = sample_uniform_cone(u, v, cos_angle);
        const Vec3 ndir = mat3x3_mul(frame, smp.dir);
        const float inv_pdf = 2 * flt_pi * (1 - cos_angle);
        return DirectLightSample{v3(0, 0, 0), neg(ndir), cmulf(col(l.p[4], l.p[5], l.p[6]), inv_pdf), Pdf{smp.pdf, PDF_SOLID}, smp.dir.z, std::numeric_limits<float>::infinity()};
    }
    case LIGHT_DIRECTIONAL: {  // light/directional.art:6
        const float scene_radius = len(sc.bbox.max - sc.bbox.min) / 2 * 1.01f;
        const Vec3 dir = v3(l.p[0], l.p[1], l.p[2]);
        return DirectLightSample{from.point + mulf(dir, -scene_radius), neg(dir), col(l.p[3], l.p[4], l.p[5]), Pdf{1, PDF_DELTA}, 1, scene_radius};
    }
    case LIGHT_POINT: {  // light/point.art:3-8
        const Vec3 pos = v3(l.p[0], l.p[1], l.p[2]);
        const Vec3 dir_ = pos - from.point;
        const float dist = len(dir_);
        const Vec3 dir = mulf(dir_, safe_div(1, dist));
        return DirectLightSample{pos, dir, col(l.p[3], l.p[4], l.p[5]), Pdf{1, PDF_AREA}, 1, dist};
    }
    case LIGHT_SPOT: {  // light/spot.art:8-44; p = pos, dir, cos(cutoff), cos(falloff), intensity (the cosines are host constants)
        const Vec3 pos = v3(l.p[0], l.p[1], l.p[2]), sdir = v3(l.p[3], l.p[4], l.p[5]);
        const float cos_cutoff = l.p[6], cos_falloff = l.p[7];
        const float blend = cos_falloff - cos_cutoff;
        const Vec3 out_dir_ = pos - from.point;
        const float dist = len(out_dir_);
        const Vec3 out_dir = mulf(out_dir_, safe_div(1, dist));
        const float cos_angle = dot(neg(out_dir), sdir);
        float factor;
        if (blend <= flt_eps) factor = cos_angle <= cos_cutoff ? 0.0f : 1.0f;
        else { const float x = clampf((cos_angle - cos_cutoff) / blend, 0, 1); factor = x * x * (3 - 2 * x); }   // core/common.art:241
        const float cos = -dot(out_dir, sdir);
        const float area_pdf = dot(neg(out_dir), sdir) > cos_cutoff ? 1.0f : 0.0f;
        return DirectLightSample{pos, out_dir, cmulf(col(l.p[8], l.p[9], l.p[10]), factor), Pdf{area_pdf, PDF_AREA}, cos, dist};
    }
    default: {  // light/area.art:12-25
        const float u = rnd.next_f32(); const float v = rnd.next_f32();
        SurfaceElement to; Pdf pdf; float weight; Color radiance;
        if (l.type == LIGHT_PLANE_AREA) {
            PlaneEmitter(l).sample_direct(Vec2{u, v}, from.point, to, pdf, weight);
            radiance = col(l.p[21], l.p[22], l.p[23]);
        } else if (l.type == LIGHT_SPHERE_AREA) {
            sphere_emitter_sample(sc, l, Vec2{u, v}, from.point, to, pdf, weight);
            radiance = col(l.p[0], l.p[1], l.p[2]);
        } else {
            float pdfv;
            shape_emitter_sample(sc, l.entity_id, Vec2{u, v}, to, pdfv, weight);
            pdf = Pdf{pdfv, PDF_AREA};
            radiance = col(l.p[0], l.p[1], l.p[2]);
        }
        const Vec3 dir_ = to.point - from.point;
        const float dist = len(dir_);
        const Vec3 dir = mulf(dir_, safe_div(1, dist));
        const float cos = dot(dir, to.face_normal) * (from.is_entering ? -1.0f : 1.0f);
        return DirectLightSample{to.point, dir, cmulf(radiance, weight), pdf, cos, dist};
    }
    }
}

// ------------------------------------------------------------------------------------------ material
struct Material { int id; Bsdf bsdf; const LightDesc* light; bool is_emissive; };
struct ShadingContext { int pixel; Ray ray; Hit hit; SurfaceElement surf; };

// The BSDF shader of one material at one hit: what the generated `bsdf_<id> : BSDFShader = @|ctx| ...` evaluates (HitShader.cpp:16-53,
// DiffuseBSDF.cpp:13-27, DielectricBSDF.cpp:13-41, ConductorBSDF.cpp:13-35, MapBSDF.cpp:17-55): colour parameters that are textures are
// looked up at ctx.uvw = (surf.tex_coords, 0); a bump / normal map replaces the frame the inner BSDF is built on (bsdf/map.art:39-68).
inline Bsdf make_bsdf(const Scene& sc, const MaterialDesc& md, const ShadingContext& ctx) {
    const TextureSet ts{sc.textures.data(), sc.images.data()};
    const Vec2 uv = ctx.surf.tex_coords;
    Bsdf b{};
    b.type = md.bsdf; b.local = ctx.surf.local; b.is_entering = ctx.surf.is_entering;
    if (md.map_kind == MAP_BUMP) {          // texture_dx / texture_dy, texture/common.art:28-38: forward differences with delta = 0.001
        const float delta = 0.001f;
        const Color c = eval_texture(ts, md.map_tex, uv);
        const float dx = ((eval_texture(ts, md.map_tex, Vec2{uv.x + delta, uv.y}).r - c.r) * (1 / delta));
        const float dy = ((eval_texture(ts, md.map_tex, Vec2{uv.x, uv.y + delta}).r - c.r) * (1 / delta));
        const Mat3x3& l = ctx.surf.local;
        const Vec3 N = normalize(l.c2 - mulf(mulf(l.c0, dx) + mulf(l.c1, dy), md.map_strength));
        b.local = normal_set_frame(ctx.surf, ctx.ray.dir, N);
    } else if (md.map_kind == MAP_NORMAL) { // bsdf/map.art:56-60
        const Color c = eval_texture(ts, md.map_tex, uv);
        const Mat3x3& l = ctx.surf.local;
        const Vec3 oN = to_local(l, normalize(v3(2 * c.r - 1, 2 * c.g - 1, 2 * c.b - 1)));   // mat3x3_left_mul, as the reference writes it
        const Vec3 N = md.map_strength != 1 ? normalize(l.c2 + mulf(oN - l.c2, md.map_strength)) : oN;
        b.local = normal_set_frame(ctx.surf, ctx.ray.dir, N);
    }
    auto colour = [&](int slot, const float* c) { return md.tex[slot] >= 0 ? eval_texture(ts, md.tex[slot], uv) : col(c[0], c[1], c[2]); };
    if (md.bsdf == BSDF_DIFFUSE) b.kd = colour(0, md.p);
    else if (md.bsdf == BSDF_DIELECTRIC) { b.n1 = md.p[0]; b.n2 = md.p[1]; b.ks = colour(0, md.p + 2); b.kt = colour(1, md.p + 5); }
    else {                                   // p = eta rgb, k rgb, ks rgb, mirror flag
        b.c_eta = col(md.p[0], md.p[1], md.p[2]); b.c_k = col(md.p[3], md.p[4], md.p[5]);
        b.ks = colour(0, md.p + 6); b.mirror = md.p[9] != 0.0f;
        b.alpha_u = md.alpha_u; b.alpha_v = md.alpha_v;
        b.rough = md.distribution == MICROFACET_VNDF_GGX && !check_if_delta_distribution(md.alpha_u, md.alpha_v);
    }
    return b;
}

// ------------------------------------------------------------------------------------------ path technique (technique/pathtracer.art)
struct PTRayPayload { float inv_pdf; Color contrib; int depth; float eta; };
struct Payload { float v[6]; };
inline PTRayPayload unwrap(const Payload& p) { return PTRayPayload{p.v[0], col(p.v[1], p.v[2], p.v[3]), (int)p.v[4], p.v[5]}; }   // :24-29
inline void wrap(Payload& p, const PTRayPayload& pt) { p.v[0] = pt.inv_pdf; p.v[1] = pt.contrib.r; p.v[2] = pt.contrib.g; p.v[3] = pt.contrib.b; p.v[4] = (float)pt.depth; p.v[5] = pt.eta; }  // :15-22

struct PathTracer {
    const Scene& sc;
    int max_path_len, min_path_len; float clamp_value; bool enable_nee;
    int n_inf, n_fin, num_lights; float pdf_lights;
    int selector;   // 0 uniform, 1 cdf, 2 hierarchy
    explicit PathTracer(const Scene& s) : sc(s) {
        selector = s.technique.light_selector;
        max_path_len = s.technique.max_depth; min_path_len = s.technique.min_depth; clamp_value = s.technique.clamp; enable_nee = s.technique.nee != 0;
        n_inf = (int)s.inf_lights.size(); n_fin = (int)s.fin_lights.size(); num_lights = n_inf + n_fin;
        pdf_lights = num_lights == 0 ? 1.0f : 1 / (float)num_lights;   // light/light_selector.art:26-44
    }
    Color handle_color(Color c) const { return clamp_value > 0 ? csaturate(c, clamp_value) : c; }   // :46-50
    // ---- light selectors (light/light_selector.art). `from` is the point light is gathered at.
    // cdf ("simple") selector: core/cdf.art:43-48,70-73, core/interval.art:7-23 over light_cdf.bin = [x1 .. x(n-1), 1]
    float cdf_get(int i) const { return i == 0 ? 0.0f : sc.selector_data[(size_t)i - 1]; }
    float cdf_pdf_discrete(int x) const { return cdf_get(x + 1) - cdf_get(x); }
    int cdf_sample_discrete(float u, float& pdf) const {
        const int size = n_fin + 1;
        int first = 0, len = size;
        while (len > 0) {
            const int half = len / 2, middle = first + half;
            if (cdf_get(middle) <= u) { first = middle + 1; len -= half + 1; } else len = half;
        }
        const int off = std::min(std::min(std::max(first - 1, 0), size - 1), n_fin - 1);
        pdf = cdf_pdf_discrete(off);
        return off;
    }
    // hierarchy: light/light_hierarchy.art:13-105 over light_hierarchy.bin = codes[round_up(n, 4)] then 8 words per node
    struct HEntry { Vec3 pos, dir; float flux; int id; bool has_dir, is_leaf; };
    HEntry h_load(int id) const {
        const float* e = sc.selector_data.data() + (size_t)((n_fin + 3) / 4 * 4) + (size_t)id * 8;
        int32_t index; std::memcpy(&index, e + 7, 4);
        HEntry h;
        h.pos = v3(e[0], e[1], e[2]); h.dir = v3(e[4], e[5], e[6]);
        h.flux = std::fabs(e[3]); h.id = index < 0 ? -index - 1 : index;
        h.has_dir = !std::signbit(e[3]); h.is_leaf = index >= 0;
        return h;
    }
    static float h_cost(const HEntry& e, Vec3 pos) {
        const Vec3 cdir = e.pos - pos;
        const float dist2 = len2(cdir);
        const float cos_d = e.has_dir ? std::fabs(dot(e.dir, normalize(cdir))) : 1.0f;
        return safe_div(e.flux * cos_d, dist2);
    }
    static float h_left_prop(const HEntry& l, const HEntry& r, Vec3 pos) { const float cl = h_cost(l, pos), cr = h_cost(r, pos); return 1 / (1 + cr / cl); }
    int h_sample(Rng& rnd, Vec3 pos, float& pdf) const {
        pdf = 1.0f;
        HEntry entry = h_load(0);
        while (!entry.is_leaf) {
            const HEntry left = h_load(entry.id), right = h_load(entry.id + 1);
            const float prop = h_left_prop(left, right, pos);
            const bool is_left = rnd.next_f32() < prop;
            entry = is_left ? left : right;
            pdf *= is_left ? prop : 1 - prop;
        }
        return entry.id;
    }
    float h_pdf(int light_id, Vec3 pos) const {
        uint32_t code; std::memcpy(&code, sc.selector_data.data() + light_id, 4);
        float pdf = 1.0f;
        HEntry entry = h_load(0);
        while (!entry.is_leaf) {
            const HEntry left = h_load(entry.id), right = h_load(entry.id + 1);
            const float prop = h_left_prop(left, right, pos);
            const bool is_left = (code & 1u) == 0;
            entry = is_left ? left : right;
            pdf *= is_left ? prop : 1 - prop;
            code >>= 1;
        }
        return pdf;
    }
    int pick_light_id(Rng& rnd, int n) const { return n <= 1 ? 0 : rnd.next_i32(0, n - 1); }   // light_selector.art:18-24
    // finite part of the cdf / hierarchy selectors
    int select_finite(Rng& rnd, Vec3 from, float& pdf) const {
        if (selector == 1) return cdf_sample_discrete(rnd.next_f32(), pdf);
        if (n_fin == 1) { pdf = 1.0f; return 0; }                                               // light_hierarchy.art:109-113
        return h_sample(rnd, from, pdf);
    }
    float pdf_finite(int light_id, Vec3 from) const {
        if (selector == 1) return cdf_pdf_discrete(light_id);
        return n_fin == 1 ? 1.0f : h_pdf(light_id, from);
    }
    const LightDesc& select_light(Rng& rnd, Vec3 from, float& pdf) const {
        if (selector == 0) {                                                                     // light_selector.art:26-44
            const int id = pick_light_id(rnd, num_lights);
            pdf = pdf_lights;
            return id < n_inf ? sc.inf_lights[id] : sc.fin_lights[id - n_inf];
        }
        if (n_inf == 0) return sc.fin_lights[select_finite(rnd, from, pdf)];                    // :48-55, :82-90
        const float q = rnd.next_f32();                                                          // :57-75, :92-108: half the samples go to the infinite lights
        if (q < 0.5f) { const int id = pick_light_id(rnd, n_inf); pdf = 1 / (float)n_inf * 0.5f; return sc.inf_lights[id]; }
        const int id = select_finite(rnd, from, pdf);
        pdf = pdf * (1 - 0.5f);
        return sc.fin_lights[id];
    }
    // probability of having picked this light from `from`
    float select_pdf(bool infinite, int light_id, Vec3 from) const {
        if (selector == 0) return pdf_lights;
        if (infinite) return 1 / (float)n_inf * 0.5f;
        return n_inf == 0 ? pdf_finite(light_id, from) : pdf_finite(light_id, from) * (1 - 0.5f);
    }
    // :52-117
    bool on_shadow(const ShadingContext& ctx, Rng& rnd, const Payload& payload, const Material& mat, Ray& out_ray, Color& out_color) const {
        if (!enable_nee) return false;
        if (mat.bsdf.is_all_delta() || num_lights == 0) return false;
        const PTRayPayload pt = unwrap(payload);
        if (pt.depth + 1 > max_path_len) return false;
        float light_select_pdf;
        const LightDesc& light = select_light(rnd, ctx.surf.point, light_select_pdf);
        const DirectLightSample ls = light_sample_direct(sc, light, rnd, ctx.surf);
        const float pdf_l_s = ls.pdf.as_solid(ls.cos, ls.dist * ls.dist) * light_select_pdf;
        if (pdf_l_s <= flt_eps) return false;
        const Vec3 in_dir = ls.dir;
        const Vec3 out_dir = neg(ctx.ray.dir);
        if (ls.cos > flt_eps) {
            float mis;
            if (light_delta(light)) mis = 1.0f;
            else { const float pdf_e_s = mat.bsdf.pdf(in_dir, out_dir); mis = 1 / (1 + pdf_e_s / pdf_l_s); }
            const float factor = ls.pdf.value / pdf_l_s;
            const Color contrib = handle_color(cmulf(cmul(ls.intensity, cmul(pt.contrib, mat.bsdf.eval(in_dir, out_dir))), mis * factor));
            if (caverage(contrib) <= flt_eps) return false;
            const float offset = 0.001f;
            if (light_infinite(light)) out_ray = make_ray(ctx.surf.point, in_dir, offset, flt_max, ray_flag_shadow);
            else out_ray = make_ray(ctx.surf.point, ls.pos - ctx.surf.point, offset, 1 - offset, ray_flag_shadow);
            out_color = contrib;
            return true;
        }
        return false;
    }
    // Emission of an emissive material: driver/material.art:21-28, light/area.art:38-39
    // :119-139
    bool on_hit(const ShadingContext& ctx, const Payload& payload, const Material& mat, Color& out) const {
        if (mat.is_emissive && ctx.surf.is_entering) {
            const PTRayPayload pt = unwrap(payload);
            const float dt = -dot(ctx.ray.dir, ctx.surf.local.c2);
            if (dt > flt_eps) {
                const LightDesc& l = *mat.light;
                Color intensity; Pdf pdf;
                if (l.type == LIGHT_PLANE_AREA) {
                    intensity = col(l.p[21], l.p[22], l.p[23]);
                    pdf = PlaneEmitter(l).pdf_direct(ctx.ray.org);
                } else if (l.type == LIGHT_SPHERE_AREA) {   // light/area.art:301-303
                    intensity = col(l.p[0], l.p[1], l.p[2]);
                    pdf = Pdf{safe_div(1, l.p[7]), PDF_AREA};
                } else {  // shape emitter: light/area.art:77-86,105
                    intensity = col(l.p[0], l.p[1], l.p[2]);
                    SurfaceElement s; float pdfv, w;
                    shape_emitter_sample(sc, l.entity_id, ctx.surf.prim_coords, s, pdfv, w);
                    pdf = Pdf{pdfv, PDF_AREA};
                }
                const float pdf_s = pdf.as_solid(dt, ctx.hit.distance * ctx.hit.distance);
                const float mis = enable_nee ? 1 / (1 + pt.inv_pdf * select_pdf(false, (int)(mat.light - sc.fin_lights.data()), ctx.ray.org) * pdf_s) : 1.0f;
                out = handle_color(cmulf(cmul(pt.contrib, intensity), mis));
                return true;
            }
        }
        return false;
    }
    // :141-168
    bool on_miss(const Ray& ray, const Payload& payload, Color& out) const {
        int inflights = 0;
        Color color = col(0, 0, 0);
        for (int i = 0; i < n_inf; ++i) {
            const LightDesc& l = sc.inf_lights[i];
            if (light_infinite(l) && !light_delta(l)) {
                const PTRayPayload pt = unwrap(payload);
                ++inflights;
                Color emit; float pdf_s;
                if (l.type == LIGHT_SUN) {                                   // light/sun.art:33-45
                    const bool hit = sun_hit(l, ray.dir);
                    emit = hit ? col(l.p[4], l.p[5], l.p[6]) : col(0, 0, 0);
                    pdf_s = hit ? uniform_cone_pdf(l.p[3]) : 0.0f;
                } else if (l.type == LIGHT_ENV_TEXTURED) {                   // light/env.art:145-152
                    emit = env_textured_emission(l, TextureSet{sc.textures.data(), sc.images.data()}, ray.dir);
                    pdf_s = env_textured_pdf(l, sc.aux_data.data(), ray.dir);
                } else if (l.type == LIGHT_ENV_TEX) {
                    emit = env_tex_radiance(l, TextureSet{sc.textures.data(), sc.images.data()}, ray.dir);
                    pdf_s = 1 / (4 * flt_pi);
                } else {
                    emit = col(l.p[0], l.p[1], l.p[2]);                      // light/env.art:96
                    pdf_s = 1 / (4 * flt_pi);                                // light/env.art:97, sampling.art:47-51
                }
                const float mis = enable_nee ? 1 / (1 + pt.inv_pdf * select_pdf(true, i, ray.org) * pdf_s) : 1.0f;
                color = cadd(color, handle_color(cmulf(cmul(pt.contrib, emit), mis)));
            }
        }
        if (inflights > 0) { out = color; return true; }
        return false;
    }
    // :170-210
    bool on_bounce(const ShadingContext& ctx, Rng& rnd, Payload& payload, const Material& mat, Ray& out_ray) const {
        const PTRayPayload pt = unwrap(payload);
        if (pt.depth + 1 > max_path_len) return false;
        const Vec3 out_dir = neg(ctx.ray.dir);
        BsdfSample ms;
        if (!mat.bsdf.sample(rnd, out_dir, false, ms)) return false;
        if (ms.pdf <= flt_eps) return false;
        const Color contrib = cmul(pt.contrib, ms.color);
        const float rr_prob = (pt.depth + 1 > min_path_len) ? clampf(cmaxcomp(cmulf(contrib, pt.eta * pt.eta)), 0.05f, 0.95f) : 1.0f;  // :5,190
        if (rnd.next_f32() >= rr_prob) return false;
        const float inv_pdf = ms.is_delta ? 0.0f : 1 / ms.pdf;
        const Color new_contrib = cmulf(contrib, 1 / rr_prob);
        wrap(payload, PTRayPayload{inv_pdf, new_contrib, pt.depth + 1, pt.eta * ms.eta});
        out_ray = make_ray(ctx.surf.point, ms.in_dir, 0.001f, flt_max, ray_flag_bounce);
        return true;
    }
};

// ------------------------------------------------------------------------------------------ streams (driver/streams.art:1-32)
struct PrimaryStream {
    std::vector<int> id; std::vector<Ray> ray; std::vector<Hit> hit; std::vector<uint32_t> rnd; std::vector<Payload> payload;
    void resize(size_t n) { id.resize(n); ray.resize(n); hit.resize(n); rnd.resize(n); payload.resize(n); }
    void move(size_t dst, size_t src) { id[dst] = id[src]; ray[dst] = ray[src]; hit[dst] = hit[src]; rnd[dst] = rnd[src]; payload[dst] = payload[src]; }
};
struct SecondaryStream {
    std::vector<int> id; std::vector<Ray> ray; std::vector<int> mat_id; std::vector<Color> color;
    void resize(size_t n) { id.resize(n); ray.resize(n); mat_id.resize(n); color.resize(n); }
};

struct Camera {  // camera/perspective.art:2-6,29-42
    Vec3 eye; Mat3x3 view; Vec2 scale; float tmin, tmax;
    Camera(const CameraDesc& c, int w, int h) {
        eye = v3(c.eye[0], c.eye[1], c.eye[2]);
        const Vec3 dir = v3(c.dir[0], c.dir[1], c.dir[2]), up = v3(c.up[0], c.up[1], c.up[2]);
        const Vec3 right = normalize(cross(dir, up));
        view = Mat3x3{right, up, dir};
        const float aspect = c.aspect > 0 ? c.aspect : (float)w / (float)h;
        if (c.fov_vertical) { const float sh = tanf(c.fov / 2); scale = Vec2{sh * aspect, sh}; }
        else { const float sw = tanf(c.fov / 2); scale = Vec2{sw, sw / aspect}; }
        tmin = c.tmin; tmax = c.tmax;
    }
    Ray generate_ray(float nx, float ny) const {
        const Vec3 d = normalize(mat3x3_mul(view, v3(scale.x * nx, scale.y * ny, 1)));
        return make_ray(eye, d, tmin, tmax, ray_flag_camera);
    }
};

struct Oracle {
    Scene scene;
    float* aov_normals = nullptr; float* aov_albedo = nullptr;   // standard AOVs (technique/internal/infobuffer.art), set by igo_set_aovs
    bool deterministic = false;   // igo_set_deterministic: per-sample accumulation (the device's option "deterministic")
    explicit Oracle(const SceneDesc& d) : scene(d) {}
};

// One tile: driver/mapping_cpu.art:719-861
void trace_tile(const Scene& sc, const Settings& st, const StreamRay* list_rays, int xmin, int ymin, int xmax, int ymax,
                float* fb, int use_bvh, uint64_t counters[3], float* aov_normals = nullptr, float* aov_albedo = nullptr, bool deterministic = false) {
    const int spi = st.spi;
    const int capacity = spi * 16 * 16;                         // :717
    const int W = st.width, H = st.height;
    PrimaryStream primary, tmp; SecondaryStream secondary;
    primary.resize(capacity); tmp.resize(capacity); secondary.resize(capacity);
    const PathTracer tech(sc);
    const Camera camera(sc.camera, W, H);
    const float inv_spi = 1 / (float)spi;                        // driver/accumulator.art:23-31
    const int tile_w = xmax - xmin, tile_h = ymax - ymin;
    const int num_rays = spi * tile_w * tile_h;
    // Deterministic accumulation (the device's option "deterministic"): the order in which the contributions of ONE pixel are added is
    // what makes the frame depend on scheduling -- in the reference it is the order of the tile's wavefront, here and on the device it
    // is fixed: every sample (ray id) sums its own contributions in path order, then the samples are added to the pixel in sample order.
    std::vector<float> slots, nslots, aslots;   // colour, Normals, Albedo: one rgb slot per sample of the tile
    if (deterministic) { slots.assign((size_t)num_rays * 3, 0.0f); if (aov_normals) { nslots.assign((size_t)num_rays * 3, 0.0f); aslots.assign((size_t)num_rays * 3, 0.0f); } }
    auto slot_of = [&](int ray_id) {
        const int pixel = ray_id / spi, sample = ray_id - pixel * spi;
        const int px = pixel % W - xmin, py = pixel / W - ymin;
        return ((size_t)(py * tile_w + px) * spi + sample) * 3;
    };
    auto splat = [&](int ray_id, Color c) {                      // :437-451
        float* p = deterministic ? slots.data() + slot_of(ray_id) : fb + (size_t)(ray_id / spi) * 3;
        p[0] += c.r * inv_spi; p[1] += c.g * inv_spi; p[2] += c.b * inv_spi;
    };
    const int n_ent = (int)sc.entities.size();
    std::vector<int> ray_begins(n_ent + 2), ray_ends(n_ent + 2);
    int id = 0, current_size = 0;
    while (id < num_rays || current_size > 0) {
        // ---- (re-)generate: cpu_generate_rays :313-360, make_camera_emitter driver/emitter.art:6-16
        if (current_size < capacity && id < num_rays) {
            const int n_new = std::min(num_rays - id, capacity - current_size);
            for (int i = 0; i < n_new; ++i) {
                const int in_tile_id = id + i;
                const int sample = in_tile_id % spi;
                const int in_tile_pixel = in_tile_id / spi;
                const int ty = in_tile_pixel / tile_w, tx = in_tile_pixel - ty * tile_w;
                const int x = xmin + tx, y = ymin + ty;
                const int cur = current_size + i;
                Rng rnd{create_random_seed(sample, st.iter, st.frame, x, y, st.seed), 1};
                Ray ray;
                if (list_rays) {   // make_list_emitter driver/emitter.art:18-31
                    const int lin = y * W + x;
                    StreamRay sr{{0, 0, 0}, {0, 0, 1}, 0, 0};
                    if (lin < W) sr = list_rays[lin];
                    ray = make_ray(v3(sr.org[0], sr.org[1], sr.org[2]), v3(sr.dir[0], sr.dir[1], sr.dir[2]), sr.tmin, sr.tmax, 0);
                } else {
                    const float rx = rnd.next_f32(); const float ry = rnd.next_f32();          // sampler/pixel_sampler.art:4-10
                    const float nx = 2 * ((float)x + rx) / (float)W - 1;                         // driver/camera.art:21-29
                    const float ny = 1 - 2 * ((float)y + ry) / (float)H;
                    ray = camera.generate_ray(nx, ny);
                }
                wrap(primary.payload[cur], PTRayPayload{0, col(1, 1, 1), 1, 1});                  // pathtracer.art:33-38
                primary.ray[cur] = ray;
                primary.id[cur] = (y * W + x) * spi + sample;
                primary.rnd[cur] = rnd.counter;
            }
            current_size += n_new; id += n_new;
            counters[0] += (uint64_t)n_new;
        }
        if (n_ent == 0) {  // :759-761
            for (int i = 0; i < current_size; ++i) { Color c; if (tech.on_miss(primary.ray[i], primary.payload[i], c)) splat(primary.id[i], c); }
            current_size = 0;
            continue;
        }
        // ---- traverse primary :764
        for (int i = 0; i < current_size; ++i) primary.hit[i] = traverse(sc, primary.ray[i], false, use_bvh);
        // ---- sort by entity :63-103 (stable counting sort; misses last)
        std::fill(ray_ends.begin(), ray_ends.end(), 0);
        auto bin = [&](int i) { const int k = primary.hit[i].ent_id; return k == -1 ? n_ent : k; };
        for (int i = 0; i < current_size; ++i) ray_ends[bin(i)]++;
        int n = 0;
        for (int i = 0; i <= n_ent; ++i) { ray_begins[i] = n; n += ray_ends[i]; ray_ends[i] = n; }
        {
            std::vector<int> cursor(ray_begins.begin(), ray_begins.begin() + n_ent + 1);
            for (int i = 0; i < current_size; ++i) { const int k = cursor[bin(i)]++; tmp.id[k] = primary.id[i]; tmp.ray[k] = primary.ray[i]; tmp.hit[k] = primary.hit[i]; tmp.rnd[k] = primary.rnd[i]; tmp.payload[k] = primary.payload[i]; }
            std::swap(primary, tmp);
        }
        const int total = current_size;
        current_size = ray_ends[n_ent - 1];   // hits only
        // ---- hit shading per entity :773-784, cpu_hit_shade :467-559
        int begin = 0, ent_id = 0;
        for (int mat_id = 0; mat_id < sc.num_materials; ++mat_id) {
            for (int k = 0; k < sc.entity_per_material[mat_id]; ++k) {
                const int end = ray_ends[ent_id++];
                for (int i = begin; i < end; ++i) {
                    const Ray ray = primary.ray[i];
                    const Hit hit = primary.hit[i];
                    const int ray_id = primary.id[i];
                    const int sample = ray_id % spi, pixel_l = ray_id / spi;
                    const int px = pixel_l % W, py = pixel_l / W;
                    Rng rnd{create_random_seed(sample, st.iter, st.frame, px, py, st.seed), primary.rnd[i]};
                    const Entity& entity = sc.entities[hit.ent_id];
                    const Shape& shape = sc.shapes[entity.shape_id];
                    ShadingContext ctx;
                    ctx.pixel = pixel_l; ctx.ray = ray; ctx.hit = hit;
                    ctx.surf = shape.type == SHAPE_TRIMESH ? trimesh_surface_element(shape.mesh, entity, ray, hit)
                                                           : sphere_surface_element(shape.sph_origin, entity, ray, hit);
                    // material shader: runtime/shader/HitShader.cpp:16-53, driver/material.art
                    const MaterialDesc& md = sc.materials[mat_id];
                    Material mat;
                    mat.id = mat_id;
                    mat.bsdf = make_bsdf(sc, md, ctx);
                    mat.is_emissive = md.light_id >= 0;
                    mat.light = mat.is_emissive ? &sc.fin_lights[md.light_id] : nullptr;
                    Color hc;
                    // wrap_infobuffer_renderer, technique/internal/infobuffer.art:9-24
                    if (aov_normals && st.iter == 0 && (ray.flags & ray_flag_camera) == ray_flag_camera) {
                        const Vec3 n = ctx.surf.local.c2;
                        const Color alb = mat.bsdf.albedo(neg(ray.dir));
                        float* pn = deterministic ? nslots.data() + slot_of(ray_id) : aov_normals + (size_t)(ray_id / spi) * 3;
                        float* pa = deterministic ? aslots.data() + slot_of(ray_id) : aov_albedo + (size_t)(ray_id / spi) * 3;
                        pn[0] += n.x * inv_spi; pn[1] += n.y * inv_spi; pn[2] += n.z * inv_spi;
                        pa[0] += fminf(alb.r, 1.0f) * inv_spi; pa[1] += fminf(alb.g, 1.0f) * inv_spi; pa[2] += fminf(alb.b, 1.0f) * inv_spi;
                    }
                    if (tech.on_hit(ctx, primary.payload[i], mat, hc)) splat(ray_id, hc);
                    Ray sray; Color scol;
                    if (tech.on_shadow(ctx, rnd, primary.payload[i], mat, sray, scol)) {
                        secondary.ray[i] = sray; secondary.mat_id[i] = mat.id + 1; secondary.color[i] = scol; secondary.id[i] = ray_id;
                    } else secondary.id[i] = -1;
                    Ray nray;
                    if (tech.on_bounce(ctx, rnd, primary.payload[i], mat, nray)) { primary.ray[i] = nray; primary.rnd[i] = rnd.counter; }
                    else primary.id[i] = -1;
                }
                begin = end;
            }
        }
        // ---- miss shading :787-790, cpu_miss_shade :582-617
        for (int i = begin; i < total; ++i) { Color c; if (tech.on_miss(primary.ray[i], primary.payload[i], c)) splat(primary.id[i], c); primary.id[i] = -1; }
        // ---- compaction :794-798 (cpu_compact_* :205-310)
        int secondary_size = current_size;
        { int k = 0; for (int i = 0; i < current_size; ++i) if (primary.id[i] >= 0) { if (k != i) primary.move(k, i); ++k; } current_size = k; }
        counters[2] += (uint64_t)current_size;
        { int k = 0; for (int i = 0; i < secondary_size; ++i) if (secondary.id[i] >= 0) { if (k != i) { secondary.id[k] = secondary.id[i]; secondary.ray[k] = secondary.ray[i]; secondary.mat_id[k] = secondary.mat_id[i]; secondary.color[k] = secondary.color[i]; } ++k; } secondary_size = k; }
        // ---- shadow rays :800-853 ; hit writer driver/streams.art:112-118 (miss <=> mat_id negative)
        if (secondary_size > 0) {
            counters[1] += (uint64_t)secondary_size;
            for (int i = 0; i < secondary_size; ++i) {
                const Hit h = traverse(sc, secondary.ray[i], true, use_bvh);
                if (h.prim_id < 0) splat(secondary.id[i], secondary.color[i]);
            }
        }
    }
    if (deterministic) {   // fold the samples into their pixels, sample 0 first (api.cu k_resolve)
        auto resolve = [&](float* dst, const std::vector<float>& src) {
            for (int py = 0; py < tile_h; ++py) for (int px = 0; px < tile_w; ++px) for (int ch = 0; ch < 3; ++ch) {
                float acc = dst[((size_t)(ymin + py) * W + xmin + px) * 3 + ch];
                for (int smp = 0; smp < spi; ++smp) acc += src[((size_t)(py * tile_w + px) * spi + smp) * 3 + ch];
                dst[((size_t)(ymin + py) * W + xmin + px) * 3 + ch] = acc;
            }
        };
        resolve(fb, slots);
        if (aov_normals && st.iter == 0) { resolve(aov_normals, nslots); resolve(aov_albedo, aslots); }
    }
}

}  // namespace

// ========================================================================================== C entry points
extern "C" {

void* igo_create(const SceneDesc* d) { return new Oracle(*d); }
void igo_destroy(void* o) { delete (Oracle*)o; }
// Enables the Normals / Albedo AOVs for subsequent igo_render calls (W*H*3 floats each, accumulated; null disables)
void igo_set_aovs(void* o, float* normals, float* albedo) { ((Oracle*)o)->aov_normals = normals; ((Oracle*)o)->aov_albedo = albedo; }
// Per-sample accumulation for subsequent igo_render calls (trace_tile): the counterpart of the device's option "deterministic"
void igo_set_deterministic(void* o, int on) { ((Oracle*)o)->deterministic = on != 0; }

// Renders one iteration (driver/mapping_cpu.art cpu_trace) over 16x16 tiles with `n_threads` workers into `fb`
// (W*H*3, accumulated). When part_world > 1 only the rank's tiles are rendered: the part_tile x part_tile block in tile column tx and
// tile row ty belongs to rank (tx + ty) mod part_world (diagonals: no rank is tied to a set of columns or rows). counters = {camera, shadow, bounce} rays.
void igo_render(void* o, const Settings* st, const StreamRay* rays, float* fb, int n_threads, int use_bvh,
                int part_rank, int part_world, int part_tile, uint64_t counters[3]) {
    const Scene& sc = ((Oracle*)o)->scene;
    const int W = st->width, H = st->height, T = 16;   // runtime/shader/ShaderUtils.cpp:37
    const int tx = (W + T - 1) / T, ty = (H + T - 1) / T;
    std::atomic<int> next{0};
    std::vector<std::vector<uint64_t>> cnt((size_t)std::max(1, n_threads), std::vector<uint64_t>(3, 0));
    auto worker = [&](int tid) {
        for (;;) {
            const int t = next.fetch_add(1);
            if (t >= tx * ty) break;
            const int x0 = (t % tx) * T, y0 = (t / tx) * T;
            if (part_world > 1) {
                const int ptx = (W + part_tile - 1) / part_tile;
                const int pidx = (y0 / part_tile) * ptx + (x0 / part_tile);
                if (((pidx % ptx) + (pidx / ptx)) % part_world != part_rank) continue;
            }
            trace_tile(sc, *st, rays, x0, y0, std::min(W, x0 + T), std::min(H, y0 + T), fb, use_bvh, cnt[tid].data(), rays ? nullptr : ((Oracle*)o)->aov_normals, rays ? nullptr : ((Oracle*)o)->aov_albedo, ((Oracle*)o)->deterministic);
        }
    };
    if (n_threads <= 1) worker(0);
    else { std::vector<std::thread> th; for (int i = 0; i < n_threads; ++i) th.emplace_back(worker, i); for (auto& t : th) t.join(); }
    for (auto& c : cnt) for (int k = 0; k < 3; ++k) counters[k] += c[k];
}

void igo_trace_closest(void* o, const StreamRay* rays, const uint32_t* flags, int64_t n, HitRecord* out, int use_bvh) {
    const Scene& sc = ((Oracle*)o)->scene;
    for (int64_t i = 0; i < n; ++i) {
        const StreamRay& r = rays[i];
        const Ray ray = make_ray(v3(r.org[0], r.org[1], r.org[2]), v3(r.dir[0], r.dir[1], r.dir[2]), r.tmin, r.tmax, flags ? flags[i] : ray_flag_camera);
        const Hit h = traverse(sc, ray, false, use_bvh);
        out[i] = HitRecord{h.ent_id, h.prim_id, h.distance, h.prim_coords.x, h.prim_coords.y};
    }
}
void igo_trace_any(void* o, const StreamRay* rays, const uint32_t* flags, int64_t n, int32_t* occluded, int use_bvh) {
    const Scene& sc = ((Oracle*)o)->scene;
    for (int64_t i = 0; i < n; ++i) {
        const StreamRay& r = rays[i];
        const Ray ray = make_ray(v3(r.org[0], r.org[1], r.org[2]), v3(r.dir[0], r.dir[1], r.dir[2]), r.tmin, r.tmax, flags ? flags[i] : ray_flag_shadow);
        occluded[i] = traverse(sc, ray, true, use_bvh).prim_id >= 0;
    }
}

// Known-answer hooks (src/tests/artic/test_intersection.art)
int igo_kat_tri(const float v0[3], const float e1[3], const float e2[3], const float n[3], const float org[3], const float dir[3],
                float tmin, float tmax, int cull, float out_tuv[3]) {
    const Tri tri{v3(v0[0], v0[1], v0[2]), v3(e1[0], e1[1], e1[2]), v3(e2[0], e2[1], e2[2]), v3(n[0], n[1], n[2])};
    const Ray ray = make_ray(v3(org[0], org[1], org[2]), v3(dir[0], dir[1], dir[2]), tmin, tmax, 0);
    return intersect_ray_tri_mt(cull != 0, ray, tri, out_tuv[0], out_tuv[1], out_tuv[2]) ? 1 : 0;
}
int igo_kat_box(const float bmin[3], const float bmax[3], const float org[3], const float dir[3], float tmin, float tmax, float out[2]) {
    const Ray ray = make_ray(v3(org[0], org[1], org[2]), v3(dir[0], dir[1], dir[2]), tmin, tmax, 0);
    intersect_ray_box(ray, BBox{v3(bmin[0], bmin[1], bmin[2]), v3(bmax[0], bmax[1], bmax[2])}, ray.tmax, out[0], out[1]);
    return out[0] <= out[1] ? 1 : 0;   // traversal/mapping_cpu.art:457 (miss <=> exit < entry)
}
uint32_t igo_random_seed(int sample, int iter, int frame, int x, int y, int user) { return create_random_seed(sample, iter, frame, x, y, user); }
uint32_t igo_tea(uint32_t v0, uint32_t v1) { return sample_tea_u32(v0, v1); }
float igo_next_f32(uint32_t seed, uint32_t counter) { Rng r{seed, counter}; return r.next_f32(); }
void igo_detmath(int fn, const float* a, const float* b, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        switch (fn) {
        case 0: out[i] = dm_sinf(a[i]); break;
        case 1: out[i] = dm_cosf(a[i]); break;
        case 2: out[i] = dm_acosf(a[i]); break;
        case 3: out[i] = dm_atan2f(a[i], b[i]); break;
        default: out[i] = 0;
        }
    }
}
// One sample of the pure dielectric BSDF (bsdf/dielectric.art:15-37) on a surface with normal `n` (unit), seen from out_dir
// (unit, pointing away from the surface); `entering`: the ray comes from the n1 side. seed/counter select the random number.
// out = in_dir xyz, eta, colour r (ks = 0.25, kt = 0.5 so that the branch taken is visible), Fresnel factor.
void igo_dielectric_sample(float n1, float n2, const float n[3], const float out_dir[3], int entering, uint32_t seed, uint32_t counter, float out[6]) {
    SurfaceElement surf{};
    surf.is_entering = entering != 0;
    surf.local = make_orthonormal(v3(n[0], n[1], n[2]));
    Bsdf b{}; b.type = BSDF_DIELECTRIC; b.local = surf.local; b.is_entering = surf.is_entering; b.n1 = n1; b.n2 = n2; b.ks = col(0.25f, 0.25f, 0.25f); b.kt = col(0.5f, 0.5f, 0.5f);
    Rng rnd{seed, counter};
    BsdfSample sm{};
    b.sample(rnd, v3(out_dir[0], out_dir[1], out_dir[2]), false, sm);
    const float k = surf.is_entering ? n1 / n2 : n2 / n1;
    FresnelTerm ft{0, 1};
    if (!fresnel(k, dot(v3(out_dir[0], out_dir[1], out_dir[2]), surf.local.c2), ft)) ft = FresnelTerm{0, 1};
    out[0] = sm.in_dir.x; out[1] = sm.in_dir.y; out[2] = sm.in_dir.z; out[3] = sm.eta; out[4] = sm.color.r; out[5] = ft.factor;
}
void igo_cosine_hemisphere(float u, float v, float out[4]) { const DirSample d = sample_cosine_hemisphere(u, v); out[0] = d.dir.x; out[1] = d.dir.y; out[2] = d.dir.z; out[3] = d.pdf; }
// core/warp.art: 0 = square_to_concentric_disk (:2-22), 1 = dir_from_spherical (:50-57), 2 = spherical_from_dir (:44-48) -- the forward maps the path uses,
// for the bijection tests of src/tests/artic/test_warp.art (tests/test_oracle_kat.py restates the inverse maps)
void igo_warp(int fn, const float in[3], float out[3]) {
    out[0] = out[1] = out[2] = 0;
    if (fn == 0) square_to_concentric_disk(in[0], in[1], out[0], out[1]);
    else if (fn == 1) { float st, ct, sp, cp; dm_sincosf(in[0], &st, &ct); dm_sincosf(in[1], &sp, &cp); out[0] = st * cp; out[1] = st * sp; out[2] = ct; }
    else { const float theta = dm_acosf(in[2]); float phi = dm_atan2f(in[1], in[0]); if (phi < 0) phi = phi + 2 * flt_pi; out[0] = theta; out[1] = phi; }
}
void igo_equal_area_sphere(float u, float v, float out[3]) { const Vec3 d = equal_area_square_to_sphere(u, v); out[0] = d.x; out[1] = d.y; out[2] = d.z; }
// ---- known-answer hooks for the round-2 additions (tests/test_oracle_kat.py)
// 1-D cdf over `data` = [x1 .. xn] (leading 0 virtual), src/tests/artic/test_cdf.art: out = {discrete off, discrete pdf, continuous off, pos, pdf, pdf_continuous(pos) off, pdf}
void igo_cdf1d(const float* data, int func_size, float u, float out[7]) {
    const Cdf1D c{data, func_size};
    float dpdf, pos, cpdf, ppdf;
    const int doff = c.sample_discrete(u, dpdf);
    const int coff = c.sample_continuous(u, pos, cpdf);
    const int poff = c.pdf_continuous(pos, ppdf);
    out[0] = (float)doff; out[1] = dpdf; out[2] = (float)coff; out[3] = pos; out[4] = cpdf; out[5] = (float)poff; out[6] = ppdf;
}
// 2-D cdf (make_cdf_2d_from_buffer): out = {pos x, pos y, pdf, pdf_continuous(pos)}
void igo_cdf2d(const float* data, int size_x, int size_y, float ux, float uy, float out[4]) {
    const Cdf2D c{data, size_x, size_y};
    Vec2 pos; float pdf;
    c.sample_continuous(ux, uy, pos, pdf);
    out[0] = pos.x; out[1] = pos.y; out[2] = pdf; out[3] = c.pdf_continuous(pos);
}
// GGX microfacet functions on the identity frame (src/tests/artic/test_microfacet.art): fn 0 ndf_ggx(m), 1 g_1_smith(w), 2 pdf_vndf_ggx(w, m),
// 3 sample_vndf_ggx(seed, counter; w) -> out = normal xyz, pdf
void igo_microfacet(int fn, float alpha_u, float alpha_v, const float w[3], const float m[3], uint32_t seed, uint32_t counter, float out[4]) {
    const Mat3x3 id{v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)};
    const Vec3 W = v3(w[0], w[1], w[2]), M = v3(m[0], m[1], m[2]);
    if (fn == 0) out[0] = ndf_ggx(id, M, alpha_u, alpha_v);
    else if (fn == 1) out[0] = g_1_smith(id, W, alpha_u, alpha_v);
    else if (fn == 2) out[0] = pdf_vndf_ggx(id, W, M, alpha_u, alpha_v);
    else { Rng rnd{seed, counter}; const Vec3 n = sample_vndf_ggx(rnd, id, W, alpha_u, alpha_v); out[0] = n.x; out[1] = n.y; out[2] = n.z; out[3] = pdf_vndf_ggx(id, W, n, alpha_u, alpha_v); }
}
// One sample of the rough conductor (bsdf/conductor.art:45-141, eta = 0, k = 1, ks = 1) on the frame with normal n: out = valid, in_dir xyz, pdf, colour rgb, pdf(in, out), eval(in, out) rgb
void igo_rough_conductor_sample(float alpha_u, float alpha_v, const float n[3], const float out_dir[3], uint32_t seed, uint32_t counter, float out[12]) {
    Bsdf b{}; b.type = BSDF_CONDUCTOR; b.local = make_orthonormal(v3(n[0], n[1], n[2])); b.is_entering = true;
    b.c_eta = col(0, 0, 0); b.c_k = col(1, 1, 1); b.ks = col(1, 1, 1); b.rough = true; b.alpha_u = alpha_u; b.alpha_v = alpha_v;
    Rng rnd{seed, counter};
    BsdfSample sm{};
    const Vec3 o = v3(out_dir[0], out_dir[1], out_dir[2]);
    const bool ok = b.sample(rnd, o, false, sm);
    out[0] = ok ? 1.0f : 0.0f;
    if (!ok) return;
    out[1] = sm.in_dir.x; out[2] = sm.in_dir.y; out[3] = sm.in_dir.z; out[4] = sm.pdf; out[5] = sm.color.r; out[6] = sm.color.g; out[7] = sm.color.b;
    out[8] = b.pdf(sm.in_dir, o);
    const Color e = b.eval(sm.in_dir, o); out[9] = e.r; out[10] = e.g; out[11] = e.b;
}
// texture lookup through the scene's tables (checkerboard / image filters / borders)
void igo_eval_texture(void* o, int tex, const float* uv, int64_t n, float* out_rgb) {
    const Scene& sc = ((Oracle*)o)->scene;
    const TextureSet ts{sc.textures.data(), sc.images.data()};
    for (int64_t i = 0; i < n; ++i) { const Color c = eval_texture(ts, tex, Vec2{uv[2 * i], uv[2 * i + 1]}); out_rgb[3 * i] = c.r; out_rgb[3 * i + 1] = c.g; out_rgb[3 * i + 2] = c.b; }
}
// ensure_valid_reflection + mat3x3_align_vectors (bsdf/map.art:39-45): the frame after make_normal_set; out = 9 floats, columns
void igo_normal_set_frame(const float face_n[3], const float shading_n[3], const float ray_dir[3], const float new_n[3], float out[9]) {
    SurfaceElement s{}; s.face_normal = v3(face_n[0], face_n[1], face_n[2]); s.local = make_orthonormal(v3(shading_n[0], shading_n[1], shading_n[2]));
    const Mat3x3 m = normal_set_frame(s, v3(ray_dir[0], ray_dir[1], ray_dir[2]), v3(new_n[0], new_n[1], new_n[2]));
    const Vec3 c[3] = {m.c0, m.c1, m.c2};
    for (int k = 0; k < 3; ++k) { out[3 * k] = c[k].x; out[3 * k + 1] = c[k].y; out[3 * k + 2] = c[k].z; }
}
int igo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
