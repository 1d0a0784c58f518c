// C ABI (include/igb200.h) + global kernels + host-side wavefront loop of the B200 render device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/igb200.h"
#include "bvh8.h"
#include "kernels.cuh"

using namespace igb;

// ================================================================================================ kernels
namespace {

constexpr int BLOCK = 128;

// ---- K1: ray generation (gpu_generate_rays, driver/mapping_gpu.art:616-669)
__global__ void __launch_bounds__(BLOCK) k_generate(DevScene sc, RenderParams rp, PrimaryQueue q, int* __restrict__ q_count,
                                                    long long first, int n_new, const igb200_ray* __restrict__ list_rays) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = j < n_new;
    int x = 0, y = 0, sample = 0;
    if (valid) {
        const long long g = first + j;
        const int per_tile = rp.tile_w * rp.tile_h * rp.spi;
        const int ltile = (int)(g / per_tile);
        const int rem = (int)(g - (long long)ltile * per_tile);
        const int pix = rem / rp.spi;
        sample = rem - pix * rp.spi;
        const int gt = ltile * rp.world + rp.rank;
        x = (gt % rp.tiles_x) * rp.tile_w + pix % rp.tile_w;
        y = (gt / rp.tiles_x) * rp.tile_h + pix / rp.tile_w;
        valid = x < rp.width && y < rp.height;
    }
    const int slot = warp_append(q_count, valid);
    if (!valid) return;
    Rng rnd; rnd.seed = random_seed(sample, rp.iter, rp.frame, x, y, rp.seed); rnd.counter = 1;   // driver/emitter.art:8
    V3 org, dir; float tmin, tmax; uint32_t flags;
    if (list_rays) {  // make_list_emitter, driver/emitter.art:18-31
        const int lin = y * rp.width + x;
        if (lin < rp.width) {
            const igb200_ray r = list_rays[lin];
            org = v3(r.org[0], r.org[1], r.org[2]); dir = v3(r.dir[0], r.dir[1], r.dir[2]); tmin = r.tmin; tmax = r.tmax;
        } else { org = v3(0, 0, 0); dir = v3(0, 0, 1); tmin = 0; tmax = 0; }
        flags = 0;
    } else {
        const float rx = rnd.next_f32(); const float ry = rnd.next_f32();              // sampler/pixel_sampler.art:4-10
        const float nx = 2 * ((float)x + rx) / (float)rp.width - 1;                     // driver/camera.art:21-29
        const float ny = 1 - 2 * ((float)y + ry) / (float)rp.height;
        const V3 w = v3(sc.scale_x * nx, sc.scale_y * ny, 1);                           // camera/perspective.art:34
        const V3 d = v3(dot(v3(sc.view[0], sc.view[3], sc.view[6]), w), dot(v3(sc.view[1], sc.view[4], sc.view[7]), w), dot(v3(sc.view[2], sc.view[5], sc.view[8]), w));
        dir = normalize(d);
        org = v3(sc.eye[0], sc.eye[1], sc.eye[2]); tmin = sc.cam_tmin; tmax = sc.cam_tmax; flags = RAY_CAMERA;
    }
    q.org_tmin[slot] = make_float4(org.x, org.y, org.z, tmin);
    q.dir_tmax[slot] = make_float4(dir.x, dir.y, dir.z, tmax);
    q.state[slot] = make_uint4((uint32_t)((y * rp.width + x) * rp.spi + sample), rnd.counter, 1u, __float_as_uint(1.0f));  // pathtracer.art:33-38
    q.contrib[slot] = make_float4(1, 1, 1, 0);
    q.ent[slot] = (int)flags;
}

// ---- K2: closest-hit traversal (gpu_traverse_primary, driver/mapping_gpu.art:52-77)
__global__ void __launch_bounds__(BLOCK) k_trace_primary(DevScene sc, PrimaryQueue q, const int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const float4 o = q.org_tmin[i], d = q.dir_tmax[i];
    Ray ray; ray.org = v3(o.x, o.y, o.z); ray.dir = v3(d.x, d.y, d.z); ray.tmin = o.w; ray.tmax = d.w;
    HitR h;
    trace<false>(sc, ray, (uint32_t)q.ent[i], h);
    q.hit[i] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    q.ent[i] = h.ent;
}

// ---- K3: hit + miss shading (gpu_hit_shade / gpu_miss_shade, driver/mapping_gpu.art:123-290; pathtracer.art)
__global__ void __launch_bounds__(BLOCK) k_shade(DevScene sc, RenderParams rp, PrimaryQueue q, const int* __restrict__ in_count,
                                                 PrimaryQueue nq, int* __restrict__ next_count, ShadowQueue sq, int* __restrict__ shadow_count,
                                                 float* __restrict__ fb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < *in_count;
    bool has_shadow = false, has_bounce = false;
    V3 s_org, s_dir; float s_tmax = 0; C3 s_col; int pixel = 0;
    V3 b_org, b_dir; uint4 b_state; float4 b_contrib;
    if (valid) {
        const float4 o = q.org_tmin[i], d = q.dir_tmax[i];
        const uint4 st = q.state[i];
        const float4 pc = q.contrib[i];
        const int ent = q.ent[i];
        const V3 rorg = v3(o.x, o.y, o.z), rdir = v3(d.x, d.y, d.z);
        const int ray_id = (int)st.x;
        const int sample = ray_id % rp.spi;
        pixel = ray_id / rp.spi;
        const int depth = (int)st.z;
        const float eta = __uint_as_float(st.w);
        const C3 contrib = c3(pc.x, pc.y, pc.z);
        const float inv_pdf = pc.w;
        const int n_lights = sc.n_inf + sc.n_fin;
        const float pdf_lights = n_lights == 0 ? 1.0f : 1 / (float)n_lights;          // light_selector.art:26-29
        const bool nee = sc.nee != 0;
        if (ent < 0) {
            // ---- on_miss, pathtracer.art:141-168
            int inflights = 0; C3 color = c3(0, 0, 0);
            for (int l = 0; l < sc.n_inf; ++l) {
                const float* L = sc.inf_lights + 32 * l;
                ++inflights;
                const C3 emit = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));           // env.art:96
                const float pdf_s = 1 / (4 * IGB_FLT_PI);                                // env.art:97
                const float mis = nee ? 1 / (1 + inv_pdf * pdf_lights * pdf_s) : 1.0f;
                color = cadd(color, handle_color(sc, cmulf(cmul(contrib, emit), mis)));
            }
            if (inflights > 0) splat(fb, pixel, color, rp.inv_spi);
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
            const float4 m0 = ldg4(sc.materials + 4 * mat_id), m1 = ldg4(sc.materials + 4 * mat_id + 1), m2 = ldg4(sc.materials + 4 * mat_id + 2);
            const int bsdf = __float_as_int(m0.x);
            const int light_id = __float_as_int(m0.y);
            const V3 N = surf.local.c2;
            Rng rnd; rnd.seed = random_seed(sample, rp.iter, rp.frame, pixel % rp.width, pixel / rp.width, rp.seed); rnd.counter = st.y;

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
                    } else {
                        intensity = c3(__ldg(L + 2), __ldg(L + 3), __ldg(L + 4));
                        Surf es; float pdfv, w;
                        shape_emitter_sample(sc, __float_as_int(__ldg(L + 1)), surf.pu, surf.pv, es, pdfv, w);
                        pdf.value = pdfv; pdf.measure = 1;
                    }
                    const float pdf_s = pdf_as_solid(pdf, dt, dist * dist);
                    const float mis = nee ? 1 / (1 + inv_pdf * pdf_lights * pdf_s) : 1.0f;
                    splat(fb, pixel, handle_color(sc, cmulf(cmul(contrib, intensity), mis)), rp.inv_spi);
                }
            }
            const C3 kd = c3(m0.z, m0.w, m1.x);
            const V3 out_dir = neg(rdir);
            // ---- on_shadow, pathtracer.art:52-117
            if (nee && bsdf == 0 && n_lights != 0 && !(depth + 1 > sc.max_depth)) {
                const int id = n_lights <= 1 ? 0 : rnd.next_i32(0, n_lights - 1);            // light_selector.art:18-24
                const float* L = id < sc.n_inf ? sc.inf_lights + 32 * id : sc.fin_lights + 32 * (id - sc.n_inf);
                const int lt = __float_as_int(__ldg(L));
                const LightSample ls = light_sample_direct(sc, L, lt, rnd, surf);
                const float pdf_l_s = pdf_as_solid(ls.pdf, ls.cos, ls.dist * ls.dist) * pdf_lights;
                if (!(pdf_l_s <= IGB_FLT_EPS) && ls.cos > IGB_FLT_EPS) {
                    float mis;
                    if (lt == 1) mis = 1.0f;
                    else { const float pdf_e_s = positive_cos(ls.dir, N) / IGB_FLT_PI; mis = 1 / (1 + pdf_e_s / pdf_l_s); }
                    const float factor = ls.pdf.value / pdf_l_s;
                    const C3 ev = cmulf(kd, positive_cos(ls.dir, N) * IGB_FLT_INV_PI);     // diffuse.art:3
                    const C3 cc = handle_color(sc, cmulf(cmul(ls.intensity, cmul(contrib, ev)), mis * factor));
                    if (!((cc.r + cc.g + cc.b) / 3 <= IGB_FLT_EPS)) {
                        has_shadow = true; s_col = cc; s_org = surf.point;
                        if (lt == 0) { s_dir = ls.dir; s_tmax = IGB_FLT_MAX; }
                        else { s_dir = ls.pos - surf.point; s_tmax = 1 - 0.001f; }
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
                    in_dir = m33_mul(surf.local, ld); s_color = kd; s_eta = 1; is_delta = false;
                } else {          // dielectric.art:18-34
                    const float n1 = m0.z, n2 = m0.w;
                    const C3 ks = c3(m1.x, m1.y, m1.z), kt = c3(m1.w, m2.x, m2.y);
                    const float k = surf.is_entering ? n1 / n2 : n2 / n1;
                    const float cos_o = dot(out_dir, N);
                    float cos_t = 0, factor = 1;
                    if (!fresnel(k, cos_o, cos_t, factor)) { cos_t = 0; factor = 1; }
                    if (rnd.next_f32() > factor) { in_dir = mulf(N, k * cos_o - cos_t) - mulf(out_dir, k); s_color = kt; s_eta = k; }   // vector.art:127
                    else { in_dir = mulf(N, 2 * dot(N, out_dir)) - out_dir; s_color = ks; s_eta = 1; }                                 // vector.art:124
                    s_pdf = 1; is_delta = true;
                }
                if (!(s_pdf <= IGB_FLT_EPS)) {
                    const C3 nc = cmul(contrib, s_color);
                    const C3 sc2 = cmulf(nc, eta * eta);
                    const float rr = (depth + 1 > sc.min_depth) ? clampf(fmaxf(fmaxf(sc2.r, sc2.g), sc2.b), 0.05f, 0.95f) : 1.0f;
                    if (!(rnd.next_f32() >= rr)) {
                        const C3 fc = cmulf(nc, 1 / rr);
                        has_bounce = true;
                        b_org = surf.point; b_dir = in_dir;
                        b_state = make_uint4(st.x, rnd.counter, (uint32_t)(depth + 1), __float_as_uint(eta * s_eta));
                        b_contrib = make_float4(fc.r, fc.g, fc.b, is_delta ? 0.0f : 1 / s_pdf);
                    }
                }
            }
        }
    }
    const int ss = warp_append(shadow_count, has_shadow);
    if (has_shadow) {
        sq.org_tmin[ss] = make_float4(s_org.x, s_org.y, s_org.z, 0.001f);
        sq.dir_tmax[ss] = make_float4(s_dir.x, s_dir.y, s_dir.z, s_tmax);
        sq.color_pix[ss] = make_float4(s_col.r, s_col.g, s_col.b, __int_as_float(pixel));
    }
    const int bs = warp_append(next_count, has_bounce);
    if (has_bounce) {
        nq.org_tmin[bs] = make_float4(b_org.x, b_org.y, b_org.z, 0.001f);
        nq.dir_tmax[bs] = make_float4(b_dir.x, b_dir.y, b_dir.z, IGB_FLT_MAX);
        nq.state[bs] = b_state;
        nq.contrib[bs] = b_contrib;
        nq.ent[bs] = (int)RAY_BOUNCE;
    }
}

// ---- K4: any-hit shadow traversal with fused splat (gpu_traverse_secondary, driver/mapping_gpu.art:79-121)
__global__ void __launch_bounds__(BLOCK) k_trace_shadow(DevScene sc, ShadowQueue sq, const int* __restrict__ count, float* __restrict__ fb, float inv_spi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const float4 o = sq.org_tmin[i], d = sq.dir_tmax[i];
    Ray ray; ray.org = v3(o.x, o.y, o.z); ray.dir = v3(d.x, d.y, d.z); ray.tmin = o.w; ray.tmax = d.w;
    HitR h;
    trace<true>(sc, ray, RAY_SHADOW, h);
    if (h.prim < 0) {
        const float4 c = sq.color_pix[i];
        splat(fb, __float_as_int(c.w), c3(c.x, c.y, c.z), inv_spi);
    }
}

// ---- parity / micro-benchmark kernels over plain ray lists
__global__ void __launch_bounds__(BLOCK) k_trace_list(DevScene sc, const igb200_ray* __restrict__ rays, const uint32_t* __restrict__ flags, int n, int any,
                                                      igb200_hit* __restrict__ out, int* __restrict__ occluded) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const igb200_ray r = rays[i];
    Ray ray; ray.org = v3(r.org[0], r.org[1], r.org[2]); ray.dir = v3(r.dir[0], r.dir[1], r.dir[2]); ray.tmin = r.tmin; ray.tmax = r.tmax;
    HitR h;
    if (any) {
        trace<true>(sc, ray, flags ? flags[i] : RAY_SHADOW, h);
        if (occluded) occluded[i] = h.prim >= 0;
    } else {
        trace<false>(sc, ray, flags ? flags[i] : RAY_CAMERA, h);
        if (out) { igb200_hit o; o.ent_id = h.ent; o.prim_id = h.prim; o.t = h.t; o.u = h.u; o.v = h.v; out[i] = o; }
    }
}

__global__ void k_detmath(int fn, const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, c;
    switch (fn) {
    case 0: dm_sincosf(a[i], &s, &c); out[i] = s; break;
    case 1: dm_sincosf(a[i], &s, &c); out[i] = c; break;
    case 2: out[i] = dm_acosf(a[i]); break;
    default: out[i] = dm_atan2f(a[i], b[i]); break;
    }
}

}  // namespace

// ================================================================================================ host side
static thread_local std::string g_error;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_error = buf;
    return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(-2, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

template <class T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t alloc(size_t count) { release(); n = count; if (!count) return cudaSuccess; return cudaMalloc(&p, count * sizeof(T)); }
    cudaError_t upload(const std::vector<T>& v) { cudaError_t e = alloc(v.size()); if (e != cudaSuccess || v.empty()) return e; return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

struct QueueMem {
    DevBuf<float4> org_tmin, dir_tmax, contrib, hit; DevBuf<uint4> state; DevBuf<int> ent;
    cudaError_t alloc(size_t c) {
        cudaError_t e;
        if ((e = org_tmin.alloc(c)) != cudaSuccess) return e;
        if ((e = dir_tmax.alloc(c)) != cudaSuccess) return e;
        if ((e = contrib.alloc(c)) != cudaSuccess) return e;
        if ((e = hit.alloc(c)) != cudaSuccess) return e;
        if ((e = state.alloc(c)) != cudaSuccess) return e;
        return ent.alloc(c);
    }
    PrimaryQueue view() { PrimaryQueue q; q.org_tmin = org_tmin.p; q.dir_tmax = dir_tmax.p; q.state = state.p; q.contrib = contrib.p; q.hit = hit.p; q.ent = ent.p; return q; }
};

struct igb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool has_scene = false;
    igb200_scene_desc desc{};      // scalar members only are kept
    DevScene dev{};
    DevBuf<float4> nodes, tris, ent_leaf, ent_shade, blob, materials;
    DevBuf<int4> shape_info;
    DevBuf<float> inf_lights, fin_lights;
    // framebuffer
    int width = 0, height = 0;
    DevBuf<float> fb;
    float* host_fb = nullptr; size_t host_fb_n = 0;
    // queues
    size_t capacity = 0, want_capacity = (size_t)1 << 23;
    QueueMem qa, qb;
    DevBuf<float4> sq_org, sq_dir, sq_col;
    DevBuf<int> counters;          // [0] queue A, [1] queue B, [2] shadow
    int* host_counters = nullptr;  // pinned, 4 ints
    // partition
    int rank = 0, world = 1, tile = 32;
    // stats
    uint64_t rays[3] = {0, 0, 0};
    double render_ms = 0;
    bool profile_kernels = false;
    double k_ms[4] = {0, 0, 0, 0}; uint64_t k_launch[4] = {0, 0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk[2] = {nullptr, nullptr};
    DevBuf<igb200_ray> list_rays;
};

static int ensure_queues(igb200_ctx* c, size_t need) {
    size_t cap = std::min(c->want_capacity, std::max<size_t>(need, 1024));
    cap = (cap + 1023) / 1024 * 1024;
    if (cap <= c->capacity) return 0;
    CU(c->qa.alloc(cap)); CU(c->qb.alloc(cap));
    CU(c->sq_org.alloc(cap)); CU(c->sq_dir.alloc(cap)); CU(c->sq_col.alloc(cap));
    c->capacity = cap;
    return 0;
}

extern "C" {

const char* igb200_last_error(void) { return g_error.c_str(); }

int igb200_version(int* major, int* minor) {
    if (major) *major = IGB200_VERSION_MAJOR;
    if (minor) *minor = IGB200_VERSION_MINOR;
    return 0;
}

int igb200_create(int cuda_device, igb200_ctx** out) {
    if (!out) return fail(-1, "igb200_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(-3, "igb200_create: no CUDA device available (%s) -- this library has no CPU fallback", cudaGetErrorString(e));
    if (cuda_device < 0 || cuda_device >= n) return fail(-1, "igb200_create: device %d out of range (%d devices)", cuda_device, n);
    CU(cudaSetDevice(cuda_device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cuda_device));
    if (prop.major != 10) return fail(-3, "igb200_create: device %d is sm_%d%d; this library is built for sm_100a only", cuda_device, prop.major, prop.minor);
    igb200_ctx* c = new igb200_ctx();
    c->device = cuda_device;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(c->counters.alloc(8));
    CU(cudaMemset(c->counters.p, 0, 8 * sizeof(int)));
    CU(cudaMallocHost(&c->host_counters, 8 * sizeof(int)));
    CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->evk[0])); CU(cudaEventCreate(&c->evk[1]));
    *out = c;
    return 0;
}

int igb200_destroy(igb200_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->host_fb) cudaFreeHost(c->host_fb);
    if (c->host_counters) cudaFreeHost(c->host_counters);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->evk[0]); cudaEventDestroy(c->evk[1]);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int igb200_set_option(igb200_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return fail(-1, "igb200_set_option: null argument");
    if (!strcmp(name, "capacity")) { if (value < 1024) return fail(-1, "capacity must be >= 1024"); c->want_capacity = (size_t)value; c->capacity = 0; return 0; }
    if (!strcmp(name, "profile_kernels")) { c->profile_kernels = value != 0; return 0; }
    return fail(-1, "igb200_set_option: unknown option '%s'", name);
}

int igb200_set_partition(igb200_ctx* c, int rank, int world, int tile_size) {
    if (!c) return fail(-1, "null context");
    if (world < 1 || rank < 0 || rank >= world || tile_size < 1) return fail(-1, "igb200_set_partition: invalid rank %d / world %d / tile %d", rank, world, tile_size);
    c->rank = rank; c->world = world; c->tile = tile_size;
    return 0;
}

int igb200_set_scene(igb200_ctx* c, const igb200_scene_desc* d) {
    if (!c || !d) return fail(-1, "igb200_set_scene: null argument");
    CU(cudaSetDevice(c->device));
    if (d->n_leaves != d->n_entities) return fail(-1, "igb200_set_scene: %d leaves for %d entities (one EntityLeaf1 per entity expected)", d->n_leaves, d->n_entities);
    if (d->shape_data_bytes % 16) return fail(-1, "igb200_set_scene: shapes dyn-table data must be a multiple of 16 bytes");
    for (int m = 0; m < d->n_materials; ++m)
        if (d->materials[m].bsdf != IGB200_BSDF_DIFFUSE && d->materials[m].bsdf != IGB200_BSDF_DIELECTRIC) return fail(-4, "igb200_set_scene: material %d has unsupported bsdf %d", m, d->materials[m].bsdf);
    for (int l = 0; l < d->n_infinite; ++l)
        if (d->infinite_lights[l].type != IGB200_LIGHT_ENV_CONST) return fail(-4, "igb200_set_scene: infinite light %d has unsupported type %d", l, d->infinite_lights[l].type);
    for (int l = 0; l < d->n_finite; ++l) {
        const int t = d->finite_lights[l].type;
        if (t != IGB200_LIGHT_POINT && t != IGB200_LIGHT_PLANE_AREA && t != IGB200_LIGHT_SHAPE_AREA) return fail(-4, "igb200_set_scene: finite light %d has unsupported type %d", l, t);
    }

    // ---- per-shape geometry: triangles in BVH leaf order + BVH8 (replaces the reference's pre-baked trimesh_primbvh table)
    std::vector<Node8> nodes;
    std::vector<float4> tris;
    std::vector<int4> shape_info(2 * (size_t)d->n_shapes);
    std::vector<int> shape_root(d->n_shapes, 0);
    std::vector<Box3> ent_boxes(d->n_entities);
    for (int i = 0; i < d->n_entities; ++i) {
        const igb200_entity_leaf& lf = d->leaves[i];
        for (int k = 0; k < 3; ++k) { ent_boxes[i].lo[k] = lf.min[k]; ent_boxes[i].hi[k] = lf.max[k]; }
    }
    // top-level tree first so that its root is node 1
    Bvh8 top = build_bvh8(ent_boxes, 1);
    nodes.insert(nodes.end(), top.nodes.begin(), top.nodes.end());
    for (int s = 0; s < d->n_shapes; ++s) {
        const igb200_lookup_entry& lk = d->shape_lookups[s];
        if (lk.offset % 16 || lk.offset >= d->shape_data_bytes) return fail(-1, "igb200_set_scene: shape %d has a bad dyn-table offset", s);
        const uint8_t* p = d->shape_data + lk.offset;
        const int off4 = (int)(lk.offset / 16);
        if (lk.type_id == IGB200_SHAPE_SPHERE) {
            shape_info[2 * s] = make_int4(1, off4, 0, 0);
            shape_info[2 * s + 1] = make_int4(0, 1, 0, 0);
            continue;
        }
        if (lk.type_id != IGB200_SHAPE_TRIMESH) return fail(-4, "igb200_set_scene: shape %d has unsupported provider %u", s, lk.type_id);
        const int32_t* h = reinterpret_cast<const int32_t*>(p);   // shapes/trimesh.art:77-96
        const int nf = h[0], nv = h[1], nn = h[2];
        const float* verts = reinterpret_cast<const float*>(p) + 12;
        const int32_t* inds = reinterpret_cast<const int32_t*>(verts + 4 * (size_t)nv + 4 * (size_t)nn);
        const int v_start = off4 + 3, n_start = v_start + nv, i_start = n_start + nn;
        shape_info[2 * s] = make_int4(0, v_start, n_start, i_start);
        shape_info[2 * s + 1] = make_int4((i_start + nf) * 2, nf, 0, 0);
        std::vector<Box3> boxes(nf);
        for (int t = 0; t < nf; ++t) {
            Box3 b = Box3::empty();
            for (int k = 0; k < 3; ++k) { const int vi = inds[4 * t + k]; if (vi < 0 || vi >= nv) return fail(-1, "igb200_set_scene: shape %d triangle %d has a bad index", s, t); b.extend(verts + 4 * (size_t)vi); }
            boxes[t] = b;
        }
        Bvh8 bvh = build_bvh8(boxes, 4);
        const int node_base = (int)nodes.size(), tri_base = (int)(tris.size() / 4);
        shape_root[s] = node_base + 1;
        for (Node8 n : bvh.nodes) {
            for (int k = 0; k < 8; ++k) {
                if (n.child[k] > 0) n.child[k] += node_base;
                else if (n.child[k] < 0) { const int r = -n.child[k] - 1; n.child[k] = -((((r >> 2) + tri_base) << 2 | (r & 3)) + 1); }
            }
            nodes.push_back(n);
        }
        for (int slot = 0; slot < nf; ++slot) {
            const int t = bvh.order[slot];
            const float* p0 = verts + 4 * (size_t)inds[4 * t], *p1 = verts + 4 * (size_t)inds[4 * t + 1], *p2 = verts + 4 * (size_t)inds[4 * t + 2];
            // runtime/bvh/TriBVHAdapter.h:40-61: e1 = p2 - p0, e2 = p0 - p1, n = stable normal of (e1, e2, p1 - p2)
            float e1[3], e2[3], e3[3], n[3];
            for (int k = 0; k < 3; ++k) { e1[k] = p2[k] - p0[k]; e2[k] = p0[k] - p1[k]; e3[k] = p1[k] - p2[k]; }
            const float ab_x = e1[2] * e2[1], ab_y = e1[0] * e2[2], ab_z = e1[1] * e2[0];
            const float bc_x = e2[2] * e3[1], bc_y = e2[0] * e3[2], bc_z = e2[1] * e3[0];
            const float cab[3] = {e1[1] * e2[2] - ab_x, e1[2] * e2[0] - ab_y, e1[0] * e2[1] - ab_z};
            const float cbc[3] = {e2[1] * e3[2] - bc_x, e2[2] * e3[0] - bc_y, e2[0] * e3[1] - bc_z};
            n[0] = std::fabs(ab_x) < std::fabs(bc_x) ? cab[0] : cbc[0];
            n[1] = std::fabs(ab_y) < std::fabs(bc_y) ? cab[1] : cbc[1];
            n[2] = std::fabs(ab_z) < std::fabs(bc_z) ? cab[2] : cbc[2];
            float pid; std::memcpy(&pid, &t, 4);
            tris.push_back(make_float4(p0[0], p0[1], p0[2], n[0]));
            tris.push_back(make_float4(e1[0], e1[1], e1[2], n[1]));
            tris.push_back(make_float4(e2[0], e2[1], e2[2], n[2]));
            tris.push_back(make_float4(pid, 0, 0, 0));
        }
    }
    // ---- entity records
    auto as_f = [](int32_t v) { float f; std::memcpy(&f, &v, 4); return f; };
    auto as_fu = [](uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; };
    std::vector<float4> ent_leaf(8 * (size_t)d->n_entities), ent_shade(6 * (size_t)d->n_entities);
    for (int slot = 0; slot < d->n_entities; ++slot) {
        const int i = top.order[slot];
        const igb200_entity_leaf& lf = d->leaves[i];
        const int ent = lf.entity_id & 0x7FFFFFFF;
        if (ent >= d->n_entities || lf.shape_id < 0 || lf.shape_id >= d->n_shapes) return fail(-1, "igb200_set_scene: leaf %d references a bad entity/shape", i);
        const int type = (int)d->shape_lookups[lf.shape_id].type_id;
        float4* L = &ent_leaf[8 * (size_t)slot];
        L[0] = make_float4(lf.min[0], lf.min[1], lf.min[2], as_fu(lf.flags));
        L[1] = make_float4(lf.max[0], lf.max[1], lf.max[2], as_f(type == IGB200_SHAPE_SPHERE ? 1 : 0));
        const float* m = lf.local;  // column major 3x4 -> rows
        L[2] = make_float4(m[0], m[3], m[6], m[9]);
        L[3] = make_float4(m[1], m[4], m[7], m[10]);
        L[4] = make_float4(m[2], m[5], m[8], m[11]);
        L[5] = make_float4(as_f(ent), as_f(shape_root[lf.shape_id]), 0, as_f(lf.shape_id));
        if (type == IGB200_SHAPE_SPHERE) { const float* sp = reinterpret_cast<const float*>(d->shape_data + d->shape_lookups[lf.shape_id].offset); L[6] = make_float4(sp[0], sp[1], sp[2], sp[3]); }
        else L[6] = make_float4(0, 0, 0, 0);
        L[7] = make_float4(0, 0, 0, 0);
    }
    for (int e = 0; e < d->n_entities; ++e) {
        const float* r = d->entities + 36 * (size_t)e;  // driver/entity.art:12-29
        int32_t shape_id, mat_id; std::memcpy(&shape_id, r + 33, 4); std::memcpy(&mat_id, r + 34, 4);
        if (shape_id < 0 || shape_id >= d->n_shapes || mat_id < 0 || mat_id >= d->n_materials) return fail(-1, "igb200_set_scene: entity %d references a bad shape/material", e);
        float4* E = &ent_shade[6 * (size_t)e];
        const float* g = r + 12; const float* nm = r + 24;
        E[0] = make_float4(g[0], g[3], g[6], g[9]);
        E[1] = make_float4(g[1], g[4], g[7], g[10]);
        E[2] = make_float4(g[2], g[5], g[8], g[11]);
        E[3] = make_float4(nm[0], nm[3], nm[6], as_f(shape_id));
        E[4] = make_float4(nm[1], nm[4], nm[7], as_f(mat_id));
        E[5] = make_float4(nm[2], nm[5], nm[8], 0);
    }
    std::vector<float4> node_f4(nodes.size() * 16);
    if (!nodes.empty()) std::memcpy(node_f4.data(), nodes.data(), nodes.size() * sizeof(Node8));
    std::vector<float4> blob(d->shape_data_bytes / 16);
    if (!blob.empty()) std::memcpy(blob.data(), d->shape_data, d->shape_data_bytes);
    std::vector<float4> mats(4 * (size_t)d->n_materials);
    if (d->n_materials) std::memcpy(mats.data(), d->materials, sizeof(igb200_material) * (size_t)d->n_materials);
    std::vector<float> infl(32 * (size_t)d->n_infinite), finl(32 * (size_t)d->n_finite);
    if (d->n_infinite) std::memcpy(infl.data(), d->infinite_lights, sizeof(igb200_light) * (size_t)d->n_infinite);
    if (d->n_finite) std::memcpy(finl.data(), d->finite_lights, sizeof(igb200_light) * (size_t)d->n_finite);

    CU(cudaStreamSynchronize(c->stream));
    CU(c->nodes.upload(node_f4)); CU(c->tris.upload(tris)); CU(c->ent_leaf.upload(ent_leaf)); CU(c->ent_shade.upload(ent_shade));
    CU(c->blob.upload(blob)); CU(c->shape_info.upload(shape_info)); CU(c->materials.upload(mats));
    CU(c->inf_lights.upload(infl)); CU(c->fin_lights.upload(finl));

    DevScene& s = c->dev;
    s.nodes = c->nodes.p; s.tris = c->tris.p; s.ent_leaf = c->ent_leaf.p; s.ent_shade = c->ent_shade.p; s.blob = c->blob.p;
    s.shape_info = c->shape_info.p; s.materials = c->materials.p; s.inf_lights = c->inf_lights.p; s.fin_lights = c->fin_lights.p;
    s.n_ent = d->n_entities; s.n_mat = d->n_materials; s.n_inf = d->n_infinite; s.n_fin = d->n_finite;
    {   // bbox_radius(scene_bbox) * 1.01: light/env.art:76, core/bbox.art:24 (fma dot as on the device)
        const float dx = d->bbox_max[0] - d->bbox_min[0], dy = d->bbox_max[1] - d->bbox_min[1], dz = d->bbox_max[2] - d->bbox_min[2];
        s.scene_radius = std::sqrt(std::fmaf(dx, dx, std::fmaf(dy, dy, dz * dz))) / 2 * 1.01f;
    }
    s.max_depth = d->technique.max_depth; s.min_depth = d->technique.min_depth; s.clamp_value = d->technique.clamp; s.nee = d->technique.nee;
    c->desc = *d;
    c->desc.entities = nullptr; c->desc.shape_lookups = nullptr; c->desc.shape_data = nullptr; c->desc.leaves = nullptr;
    c->desc.entity_per_material = nullptr; c->desc.materials = nullptr; c->desc.infinite_lights = nullptr; c->desc.finite_lights = nullptr;
    c->has_scene = true;
    return 0;
}

int igb200_resize(igb200_ctx* c, int width, int height) {
    if (!c) return fail(-1, "null context");
    if (width < 1 || height < 1) return fail(-1, "igb200_resize: invalid size %dx%d", width, height);
    CU(cudaSetDevice(c->device));
    if (width == c->width && height == c->height) return 0;
    CU(cudaStreamSynchronize(c->stream));
    c->width = width; c->height = height;
    const size_t n = (size_t)width * height * 3;
    CU(c->fb.alloc(n));
    CU(cudaMemset(c->fb.p, 0, n * sizeof(float)));
    if (c->host_fb) { cudaFreeHost(c->host_fb); c->host_fb = nullptr; }
    CU(cudaMallocHost(&c->host_fb, n * sizeof(float)));
    c->host_fb_n = n;
    return 0;
}

static bool is_color(const char* aov) { return !aov || !*aov || !strcmp(aov, "Color"); }

int igb200_clear(igb200_ctx* c, const char* aov) {
    if (!c) return fail(-1, "null context");
    if (!is_color(aov)) return fail(-4, "igb200_clear: AOV '%s' does not exist (only the colour framebuffer is supported)", aov);
    CU(cudaSetDevice(c->device));
    if (c->fb.p) CU(cudaMemsetAsync(c->fb.p, 0, c->fb.n * sizeof(float), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int igb200_framebuffer(igb200_ctx* c, const char* aov, float** host_ptr) {
    if (!c || !host_ptr) return fail(-1, "igb200_framebuffer: null argument");
    if (!is_color(aov)) return fail(-4, "igb200_framebuffer: AOV '%s' does not exist", aov);
    if (!c->fb.p) return fail(-1, "igb200_framebuffer: no framebuffer (call igb200_resize first)");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->host_fb, c->fb.p, c->fb.n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *host_ptr = c->host_fb;
    return 0;
}

int igb200_framebuffer_device(igb200_ctx* c, const char* aov, float** device_ptr) {
    if (!c || !device_ptr) return fail(-1, "igb200_framebuffer_device: null argument");
    if (!is_color(aov)) return fail(-4, "igb200_framebuffer_device: AOV '%s' does not exist", aov);
    if (!c->fb.p) return fail(-1, "igb200_framebuffer_device: no framebuffer");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    *device_ptr = c->fb.p;
    return 0;
}

int igb200_upload_framebuffer(igb200_ctx* c, const char* aov, const float* host_rgb) {
    if (!c || !host_rgb) return fail(-1, "igb200_upload_framebuffer: null argument");
    if (!is_color(aov)) return fail(-4, "igb200_upload_framebuffer: AOV '%s' does not exist", aov);
    if (!c->fb.p) return fail(-1, "igb200_upload_framebuffer: no framebuffer");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->fb.p, host_rgb, c->fb.n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int igb200_stats(igb200_ctx* c, uint64_t out[3], double* render_ms) {
    if (!c) return fail(-1, "null context");
    if (out) { out[0] = c->rays[0]; out[1] = c->rays[1]; out[2] = c->rays[2]; }
    if (render_ms) *render_ms = c->render_ms;
    return 0;
}

int igb200_reset_stats(igb200_ctx* c) {
    if (!c) return fail(-1, "null context");
    c->rays[0] = c->rays[1] = c->rays[2] = 0; c->render_ms = 0;
    for (int k = 0; k < 4; ++k) { c->k_ms[k] = 0; c->k_launch[k] = 0; }
    return 0;
}

int igb200_kernel_times(igb200_ctx* c, double out_ms[4], uint64_t out_launches[4]) {
    if (!c) return fail(-1, "null context");
    for (int k = 0; k < 4; ++k) { if (out_ms) out_ms[k] = c->k_ms[k]; if (out_launches) out_launches[k] = c->k_launch[k]; }
    return 0;
}

static inline int grid_for(long long n) { return (int)((n + BLOCK - 1) / BLOCK); }

int igb200_render(igb200_ctx* c, const igb200_settings* st, const igb200_ray* rays, size_t n_rays) {
    if (!c || !st) return fail(-1, "igb200_render: null argument");
    if (!c->has_scene) return fail(-1, "igb200_render: no scene assigned");
    if (st->spi < 1) return fail(-1, "igb200_render: spi must be >= 1");
    CU(cudaSetDevice(c->device));
    int W = st->width, H = st->height;
    if (rays) { W = (int)n_rays; H = 1; }   // Runtime::trace, Runtime.cpp:389-446: film = rays x 1
    if (W != c->width || H != c->height) { const int r = igb200_resize(c, W, H); if (r) return r; }

    RenderParams rp;
    rp.spi = st->spi; rp.iter = st->iter; rp.frame = st->frame; rp.seed = st->seed; rp.width = W; rp.height = H;
    rp.inv_spi = 1 / (float)st->spi;
    if (rays) { rp.tile_w = W; rp.tile_h = 1; rp.rank = 0; rp.world = 1; }
    else { rp.tile_w = c->tile; rp.tile_h = c->tile; rp.rank = c->rank; rp.world = c->world; }
    rp.tiles_x = (W + rp.tile_w - 1) / rp.tile_w;
    const int tiles_y = (H + rp.tile_h - 1) / rp.tile_h;
    const long long tiles_total = (long long)rp.tiles_x * tiles_y;
    const long long local_tiles = tiles_total > rp.rank ? (tiles_total - rp.rank + rp.world - 1) / rp.world : 0;
    const long long total = local_tiles * rp.tile_w * rp.tile_h * rp.spi;   // padded ray domain of this rank

    // camera: camera/perspective.art:2-6,29-35
    DevScene sc = c->dev;
    {
        const igb200_camera& cam = c->desc.camera;
        const float dx = cam.dir[0], dy = cam.dir[1], dz = cam.dir[2], ux = cam.up[0], uy = cam.up[1], uz = cam.up[2];
        float rx = dy * uz - dz * uy, ry = dz * ux - dx * uz, rz = dx * uy - dy * ux;
        const float rl = 1.0f / std::sqrt(std::fmaf(rx, rx, std::fmaf(ry, ry, rz * rz)));
        rx *= rl; ry *= rl; rz *= rl;
        const float v[9] = {rx, ry, rz, ux, uy, uz, dx, dy, dz};
        std::memcpy(sc.view, v, sizeof(v));
        sc.eye[0] = cam.eye[0]; sc.eye[1] = cam.eye[1]; sc.eye[2] = cam.eye[2];
        const float aspect = cam.aspect > 0 ? cam.aspect : (float)W / (float)H;
        if (cam.fov_vertical) { const float sh = tanf(cam.fov / 2); sc.scale_x = sh * aspect; sc.scale_y = sh; }
        else { const float sw = tanf(cam.fov / 2); sc.scale_x = sw; sc.scale_y = sw / aspect; }
        sc.cam_tmin = cam.tmin; sc.cam_tmax = cam.tmax;
    }
    const igb200_ray* d_rays = nullptr;
    if (rays) {
        CU(c->list_rays.alloc(n_rays));
        CU(cudaMemcpyAsync(c->list_rays.p, rays, n_rays * sizeof(igb200_ray), cudaMemcpyHostToDevice, c->stream));
        d_rays = c->list_rays.p;
    }
    { const int r = ensure_queues(c, (size_t)std::max<long long>(total, 1)); if (r) return r; }
    const long long cap = (long long)c->capacity;

    int* cnt = c->counters.p;
    CU(cudaMemsetAsync(cnt, 0, 8 * sizeof(int), c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    QueueMem* cur = &c->qa; QueueMem* nxt = &c->qb;
    int ci = 0;                       // counter index of the current queue
    long long next_id = 0, n_cur = 0; // n_cur: exact size of the current queue as known by the host
    uint64_t camera = 0, shadow = 0, bounce = 0;
    auto tick = [&](int k, bool begin) -> cudaError_t {
        if (!c->profile_kernels) return cudaSuccess;
        if (begin) return cudaEventRecord(c->evk[0], c->stream);
        cudaError_t e = cudaEventRecord(c->evk[1], c->stream);
        if (e != cudaSuccess) return e;
        e = cudaEventSynchronize(c->evk[1]);
        if (e != cudaSuccess) return e;
        float ms = 0; e = cudaEventElapsedTime(&ms, c->evk[0], c->evk[1]);
        c->k_ms[k] += ms; c->k_launch[k]++;
        return e;
    };
    while (next_id < total || n_cur > 0) {
        long long upper = n_cur;   // upper bound of the current queue size after generation
        if (next_id < total && n_cur < cap) {
            const long long n_new = std::min(total - next_id, cap - n_cur);
            CU(tick(0, true));
            k_generate<<<grid_for(n_new), BLOCK, 0, c->stream>>>(sc, rp, cur->view(), cnt + ci, next_id, (int)n_new, d_rays);
            CU(tick(0, false));
            next_id += n_new; upper += n_new;
        }
        if (sc.n_ent > 0) {
            CU(tick(1, true));
            k_trace_primary<<<grid_for(upper), BLOCK, 0, c->stream>>>(sc, cur->view(), cnt + ci);
            CU(tick(1, false));
        } else {
            // no geometry: every ray is a miss (driver/mapping_cpu.art:759-761); mark entity = -1
            CU(cudaMemsetAsync(cur->ent.p, 0xFF, (size_t)upper * sizeof(int), c->stream));
        }
        CU(cudaMemsetAsync(cnt + (1 - ci), 0, sizeof(int), c->stream));
        CU(cudaMemsetAsync(cnt + 2, 0, sizeof(int), c->stream));
        CU(tick(2, true));
        k_shade<<<grid_for(upper), BLOCK, 0, c->stream>>>(sc, rp, cur->view(), cnt + ci, nxt->view(), cnt + (1 - ci),
                                                          ShadowQueue{c->sq_org.p, c->sq_dir.p, c->sq_col.p}, cnt + 2, c->fb.p);
        CU(tick(2, false));
        CU(tick(3, true));
        k_trace_shadow<<<grid_for(upper), BLOCK, 0, c->stream>>>(sc, ShadowQueue{c->sq_org.p, c->sq_dir.p, c->sq_col.p}, cnt + 2, c->fb.p, rp.inv_spi);
        CU(tick(3, false));
        CU(cudaMemcpyAsync(c->host_counters, cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const long long gen_now = (long long)c->host_counters[ci] - n_cur;   // rays actually generated this step (valid pixels only)
        camera += (uint64_t)std::max<long long>(gen_now, 0);
        n_cur = c->host_counters[1 - ci];
        shadow += (uint64_t)c->host_counters[2];
        bounce += (uint64_t)n_cur;
        std::swap(cur, nxt); ci = 1 - ci;
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->render_ms += ms;
    c->rays[0] += camera; c->rays[1] += shadow; c->rays[2] += bounce;
    CU(cudaGetLastError());
    return 0;
}

static int trace_list(igb200_ctx* c, const igb200_ray* rays, const uint32_t* flags, size_t n, int any, igb200_hit* out, int32_t* occ) {
    if (!c || !rays || (!out && !occ)) return fail(-1, "igb200_trace_*: null argument");
    if (!c->has_scene) return fail(-1, "igb200_trace_*: no scene assigned");
    CU(cudaSetDevice(c->device));
    if (n == 0) return 0;
    DevBuf<igb200_ray> dr; DevBuf<uint32_t> df; DevBuf<igb200_hit> dh; DevBuf<int> dout;
    CU(dr.alloc(n)); CU(cudaMemcpy(dr.p, rays, n * sizeof(igb200_ray), cudaMemcpyHostToDevice));
    if (flags) { CU(df.alloc(n)); CU(cudaMemcpy(df.p, flags, n * sizeof(uint32_t), cudaMemcpyHostToDevice)); }
    if (any) CU(dout.alloc(n)); else CU(dh.alloc(n));
    k_trace_list<<<grid_for((long long)n), BLOCK, 0, c->stream>>>(c->dev, dr.p, df.p, (int)n, any, dh.p, dout.p);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (any) CU(cudaMemcpy(occ, dout.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    else CU(cudaMemcpy(out, dh.p, n * sizeof(igb200_hit), cudaMemcpyDeviceToHost));
    return 0;
}

int igb200_trace_closest(igb200_ctx* c, const igb200_ray* rays, const uint32_t* flags, size_t n, igb200_hit* out) { return trace_list(c, rays, flags, n, 0, out, nullptr); }
int igb200_trace_any(igb200_ctx* c, const igb200_ray* rays, const uint32_t* flags, size_t n, int32_t* occluded) { return trace_list(c, rays, flags, n, 1, nullptr, occluded); }

int igb200_bench_trace(igb200_ctx* c, const igb200_ray* rays, size_t n, int any_hit, int repeat, double* ms_per_pass) {
    if (!c || !rays || !ms_per_pass || repeat < 1) return fail(-1, "igb200_bench_trace: bad argument");
    if (!c->has_scene) return fail(-1, "igb200_bench_trace: no scene assigned");
    CU(cudaSetDevice(c->device));
    DevBuf<igb200_ray> dr; DevBuf<igb200_hit> dh; DevBuf<int> dout;
    CU(dr.alloc(n)); CU(cudaMemcpy(dr.p, rays, n * sizeof(igb200_ray), cudaMemcpyHostToDevice));
    if (any_hit) CU(dout.alloc(n)); else CU(dh.alloc(n));
    for (int w = 0; w < 3; ++w) k_trace_list<<<grid_for((long long)n), BLOCK, 0, c->stream>>>(c->dev, dr.p, nullptr, (int)n, any_hit, dh.p, dout.p);
    CU(cudaEventRecord(c->ev0, c->stream));
    for (int r = 0; r < repeat; ++r) k_trace_list<<<grid_for((long long)n), BLOCK, 0, c->stream>>>(c->dev, dr.p, nullptr, (int)n, any_hit, dh.p, dout.p);
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_pass = ms / repeat;
    CU(cudaGetLastError());
    return 0;
}

int igb200_test_detmath(igb200_ctx* c, int fn, const float* a, const float* b, float* out, size_t n) {
    if (!c || !a || !out) return fail(-1, "igb200_test_detmath: null argument");
    CU(cudaSetDevice(c->device));
    DevBuf<float> da, db, dout;
    CU(da.alloc(n)); CU(db.alloc(n)); CU(dout.alloc(n));
    CU(cudaMemcpy(da.p, a, n * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db.p, b ? b : a, n * sizeof(float), cudaMemcpyHostToDevice));
    k_detmath<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(fn, da.p, db.p, dout.p, (int)n);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
