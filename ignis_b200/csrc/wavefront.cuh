// The persistent wavefront kernel: the whole generate -> traverse -> shade -> next-bounce loop of one render()
// iteration in ONE cooperative launch (reference: cpu_trace / gpu_trace, driver/mapping_cpu.art:719-861,
// driver/mapping_gpu.art:727-867, and the stage kernels they launch).
//
// Loop turn (two grid barriers instead of the reference's ~10 launches + 3 host round trips per bounce):
//   phase S  shade the current primary queue -> continuation rays appended to the next primary queue, shadow rays
//            to the shadow queue (warp ballot / prefix, one atomic per warp); then REGENERATE: new camera rays
//            fill what is left of the capacity (driver/mapping_gpu.art:756-765);
//   phase T  trace: closest hit for the next primary queue and any-hit + fused framebuffer splat for the shadow
//            queue (driver/mapping_gpu.art:52-121), both kinds mixed in the same warps.
// Phase T is a persistent-threads loop: every warp fetches batches of rays from one global counter and refills
// its finished lanes with new rays as soon as fewer than `refill` lanes are walking, so SIMT utilisation does not
// collapse to the longest ray of each 32 (measured 6.7 of 32 active lanes without it, profiles/).
// Nothing is sorted or compacted in a separate pass, no count ever goes back to the host inside an iteration.
#pragma once

#include <cooperative_groups.h>

#include "../../include/igb200.h"
#include "shade.cuh"
#include "traverse.cuh"

namespace igb {

namespace cg = cooperative_groups;

constexpr int TURN_LOG = 128;

struct Control {                      // device-global; zeroed by the host before every launch
    int count[2];                     // sizes of the two primary queues
    int shadow_count[2];              // shadow-queue size, double buffered (one is reset while the other is read)
    int fetch_trace, fetch_pad[3];
    unsigned long long stat[4];       // camera, shadow, bounce rays; framebuffer splats
    unsigned long long phase_ns[2];   // time spent in trace / shade+generate phases (block 0's view, incl. barrier)
    unsigned long long phases[2];
    unsigned long long turns;
    unsigned int turn_items[TURN_LOG];   // diagnostics: rays traced in turn k ...
    unsigned int turn_trace_ns[TURN_LOG];  // ... time of its trace phase and of the shade + generate phase before it
    unsigned int turn_shade_ns[TURN_LOG];
    unsigned long long step_stats[2][8];   // diagnostics build (IGB_STEP_STATS): turns < 16 / >= 16: node, leaf, entity visits, max per ray, rays
};

struct WaveParams {
    DevScene sc;
    RenderParams rp;
    PrimaryQueue q[2];
    ShadowQueue sq;
    float* fb;
    Control* ctl;
    long long total;                  // camera-ray domain of this rank (whole tiles, may extend past the frame)
    int capacity;                     // records per queue
    const igb200_ray* list_rays;      // igtrace list emitter (driver/emitter.art:18-31) or null
    int stage_nodes, stage_tris, stage_ent;   // how many nodes / triangle slots / entity leaves go to shared memory
    int refill;                       // refill a warp when fewer lanes than this are walking
};

// ---- shared memory: [stack SMEM_STACK x blockDim uint2][nodes][tris][entity leaves], staged by TMA bulk copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_scene(const DevScene& sc, int n_nodes, int n_tris, int n_ent, unsigned char* base, Staged& sg) {
    __shared__ __align__(8) unsigned long long mbar;
    float4* s_nodes = reinterpret_cast<float4*>(base);
    float4* s_tris = s_nodes + (size_t)n_nodes * 16;
    float4* s_leaf = s_tris + (size_t)n_tris * 3;
    sg.nodes = s_nodes; sg.n_nodes = n_nodes; sg.tris = s_tris; sg.n_tris = n_tris; sg.ent_leaf = s_leaf; sg.n_ent = n_ent;
    const uint32_t bytes_n = (uint32_t)n_nodes * 256u, bytes_t = (uint32_t)n_tris * 48u, bytes_l = (uint32_t)n_ent * 128u;
    if (bytes_n + bytes_t + bytes_l == 0) return;
    const uint32_t bar = smem_u32(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_n + bytes_t + bytes_l) : "memory");
        // TMA 1-D bulk copies global -> shared (UBLKCP), in pieces of at most 32 KB
        auto bulk = [&](float4* dst, const float4* src, uint32_t bytes) {
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t n = min(32768u, bytes - off);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(reinterpret_cast<unsigned char*>(dst) + off)), "l"(reinterpret_cast<const unsigned char*>(src) + off), "r"(n), "r"(bar) : "memory");
            }
        };
        bulk(s_nodes, sc.nodes, bytes_n);
        bulk(s_tris, sc.tris, bytes_t);
        bulk(s_leaf, sc.ent_leaf, bytes_l);
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
}

// ---- phase S, part 1: hit / miss shading of records [0, n) of `q`
__device__ __forceinline__ void phase_shade(const WaveParams& P, const PrimaryQueue& q, int n, const PrimaryQueue& nq, int* next_count, int* shadow_count,
                                            int warp_id, int n_warps) {
    const int lane = threadIdx.x & 31;
    int n_splat = 0;
    ShadeSink sink; sink.nq = nq; sink.next_count = next_count; sink.sq = P.sq; sink.shadow_count = shadow_count;
    for (int base = warp_id * 32; base < n; base += n_warps * 32) {
        const int i = base + lane;
        if (i < n) n_splat += shade_record(P.sc, P.rp, q, i, P.fb, sink);
    }
    __syncwarp();
    warp_count(&P.ctl->stat[3], n_splat);
}

// ---- phase S, part 2: ray generation for domain ids [first, first + n_new) (gpu_generate_rays, driver/mapping_gpu.art:616-669)
__device__ __forceinline__ void phase_generate(const WaveParams& P, const PrimaryQueue& q, int* q_count, long long first, int n_new, int warp_id, int n_warps) {
    const int lane = threadIdx.x & 31;
    const DevScene& sc = P.sc;
    const RenderParams& rp = P.rp;
    int n_cam = 0;
    for (int base = warp_id * 32; base < n_new; base += n_warps * 32) {
        const int j = base + lane;
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
        if (!valid) continue;
        ++n_cam;
        Rng rnd; rnd.seed = random_seed(sample, rp.iter, rp.frame, x, y, rp.seed); rnd.counter = 1;   // driver/emitter.art:8
        V3 org, dir; float tmin, tmax; uint32_t flags;
        if (P.list_rays) {  // make_list_emitter, driver/emitter.art:18-31
            const int lin = y * rp.width + x;
            if (lin < rp.width) {
                const igb200_ray r = P.list_rays[lin];
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
    warp_count(&P.ctl->stat[0], n_cam);
}

// ---- phase T: items [0, n_primary) are records of `q` (closest hit), [n_primary, n_primary + n_shadow) shadow rays
__device__ __forceinline__ void phase_trace(const DevScene& sc, const Staged& sg, Stack& st, const PrimaryQueue& q, int n_primary, const ShadowQueue& sq, int n_shadow,
                                            float* __restrict__ fb, float inv_spi, int* fetch, unsigned long long* splat_stat, int batch, int refill,
                                            unsigned long long* step_stats = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_items = n_primary + n_shadow;
    int wb = 0, we = 0;          // this warp's private batch of items [wb, we)
    bool exhausted = false;      // the global counter ran past n_items
    bool active = false;
    int item = -1, n_splat = 0;
    const float4 *po = nullptr, *pd = nullptr;   // this lane's ray record
    Traversal T;

    auto finish = [&]() {
#ifdef IGB_STEP_STATS
        atomicAdd(&step_stats[0], (unsigned long long)T.n_node); atomicAdd(&step_stats[1], (unsigned long long)T.n_leaf); atomicAdd(&step_stats[2], (unsigned long long)T.n_ent);
        atomicMax(&step_stats[3], (unsigned long long)(T.n_node + T.n_leaf + T.n_ent)); atomicAdd(&step_stats[4], 1ull);
#endif
        if (item < n_primary) {
            q.hit[item] = make_float4(T.hit.t, T.hit.u, T.hit.v, __int_as_float(T.hit.prim));
            q.ent[item] = T.hit.ent;
        } else if (T.hit.prim < 0) {   // unoccluded: fused splat (driver/mapping_gpu.art:110-117)
            const float4 c = sq.color_pix[item - n_primary];
            splat(fb, __float_as_int(c.w), c3(c.x, c.y, c.z), inv_spi);
            ++n_splat;
        }
    };

    for (;;) {
        // ---- refill idle lanes
        unsigned idle = __ballot_sync(FULL, !active);
        while (idle && !exhausted) {
            if (wb >= we) {
                if (lane == 0) wb = atomicAdd(fetch, batch);
                wb = __shfl_sync(FULL, wb, 0);
                if (wb >= n_items) { exhausted = true; wb = we = 0; break; }
                we = min(wb + batch, n_items);
            }
            const int take = min(__popc(idle), we - wb);
            const int r = __popc(idle & lt_mask);
            if (!active && r < take) {
                item = wb + r;
                const bool primary = item < n_primary;
                const int j = primary ? item : item - n_primary;
                po = (primary ? q.org_tmin : sq.org_tmin) + j; pd = (primary ? q.dir_tmax : sq.dir_tmax) + j;
                if (T.begin(sc, po, pd, primary ? (uint32_t)q.ent[j] : RAY_SHADOW, !primary)) active = true; else finish();
            }
            wb += take;
            idle = __ballot_sync(FULL, !active);
        }
        unsigned act = __ballot_sync(FULL, active);
        if (act == 0) { if (exhausted) break; continue; }
        // ---- walk until too few lanes are left (then go back and refill), or to the end once no rays are left
        const int thresh = exhausted ? 1 : refill;
        do {
            if (active && T.turn(sc, sg, st, po, pd)) { finish(); active = false; }
            act = __ballot_sync(FULL, active);
        } while (__popc(act) >= thresh);
    }
    warp_count(splat_stat, n_splat);
}

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ int trace_batch(int n_items, int n_warps) {
    // small queues (the tail of deep paths) are spread over as many warps as possible, large ones amortise the atomic
    const int per_warp = (n_items + n_warps - 1) / n_warps;
    return max(32, min(128, (per_warp + 31) / 32 * 32));
}

template <int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_wavefront(const WaveParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::grid_group grid = cg::this_grid();
    uint2 lstack[LOCAL_STACK];
    Stack st; st.s = reinterpret_cast<uint2*>(smem) + threadIdx.x; st.stride = BLOCK; st.l = lstack;
    Staged sg;
    stage_scene(P.sc, P.stage_nodes, P.stage_tris, P.stage_ent, smem + (size_t)SMEM_STACK * BLOCK * sizeof(uint2), sg);

    const int n_warps = gridDim.x * (BLOCK / 32);
    const int warp_id = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const bool boss = blockIdx.x == 0 && threadIdx.x == 0;
    Control* ctl = P.ctl;

    int cur = 0, s = 0, n_cur = 0;
    long long next_cam = 0;
    unsigned long long t_mark = boss ? global_ns() : 0ull, sum_primary = 0, sum_shadow = 0, turns = 0;
    for (;;) {
        // ---- phase S: shade q[cur], regenerate into q[1 - cur]
        const int n_new = (int)min((long long)(P.capacity - n_cur), P.total - next_cam);
        if (n_cur > 0) phase_shade(P, P.q[cur], n_cur, P.q[1 - cur], &ctl->count[1 - cur], &ctl->shadow_count[s], warp_id, n_warps);
        if (n_new > 0) phase_generate(P, P.q[1 - cur], &ctl->count[1 - cur], next_cam, n_new, warp_id, n_warps);
        next_cam += n_new;
        if (boss) ctl->fetch_trace = 0;
        grid.sync();
        const int n_next = ctl->count[1 - cur], n_shadow = ctl->shadow_count[s];
        if (boss) {
            const unsigned long long t = global_ns();
            ctl->phase_ns[1] += t - t_mark; ctl->phases[1]++;
            if (turns < TURN_LOG) { ctl->turn_shade_ns[turns] = (unsigned int)(t - t_mark); ctl->turn_items[turns] = (unsigned int)(n_next + n_shadow); }
            t_mark = t;
        }
        if (n_next == 0 && n_shadow == 0 && next_cam >= P.total) break;
        // ---- phase T: trace q[1 - cur] and the shadow queue
        if (boss) { ctl->count[cur] = 0; ctl->shadow_count[1 - s] = 0; sum_primary += (unsigned long long)n_next; sum_shadow += (unsigned long long)n_shadow; }
        phase_trace(P.sc, sg, st, P.q[1 - cur], n_next, P.sq, n_shadow, P.fb, P.rp.inv_spi, &ctl->fetch_trace, &ctl->stat[3],
                    trace_batch(n_next + n_shadow, n_warps), P.refill, ctl->step_stats[turns < 16 ? 0 : 1]);
        grid.sync();
        if (boss) {
            const unsigned long long t = global_ns();
            ctl->phase_ns[0] += t - t_mark; ctl->phases[0]++;
            if (turns < TURN_LOG) ctl->turn_trace_ns[turns] = (unsigned int)(t - t_mark);
            t_mark = t;
        }
        ++turns;
        cur = 1 - cur; s = 1 - s; n_cur = n_next;
    }
    if (boss) {
        ctl->stat[1] = sum_shadow;
        ctl->stat[2] = sum_primary - ctl->stat[0];   // every primary record is a camera ray or a continuation
        ctl->turns = turns;
    }
}

// ---- stand-alone trace phase over a primary queue and / or a shadow queue (parity hooks, traversal micro-benchmark)
template <int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_trace(const DevScene sc, const PrimaryQueue q, int n_primary, const ShadowQueue sq, int n_shadow, float* fb,
                                                            int* fetch, unsigned long long* splat_stat, int stage_nodes, int stage_tris, int stage_ent, int refill) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint2 lstack[LOCAL_STACK];
    Stack st; st.s = reinterpret_cast<uint2*>(smem) + threadIdx.x; st.stride = BLOCK; st.l = lstack;
    Staged sg;
    stage_scene(sc, stage_nodes, stage_tris, stage_ent, smem + (size_t)SMEM_STACK * BLOCK * sizeof(uint2), sg);
    phase_trace(sc, sg, st, q, n_primary, sq, n_shadow, fb, 1.0f, fetch, splat_stat, trace_batch(n_primary + n_shadow, gridDim.x * (BLOCK / 32)), refill);
}

// igb200_ray list <-> queue records (hooks only)
__global__ void k_import_rays(const igb200_ray* __restrict__ rays, const uint32_t* __restrict__ flags, uint32_t default_flags, int n, PrimaryQueue q, ShadowQueue sq, int as_shadow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const igb200_ray r = rays[i];
    const float4 o = make_float4(r.org[0], r.org[1], r.org[2], r.tmin), d = make_float4(r.dir[0], r.dir[1], r.dir[2], r.tmax);
    if (as_shadow) { sq.org_tmin[i] = o; sq.dir_tmax[i] = d; sq.color_pix[i] = make_float4(1, 0, 0, __int_as_float(i)); }
    else { q.org_tmin[i] = o; q.dir_tmax[i] = d; q.ent[i] = (int)(flags ? flags[i] : default_flags); }
}
__global__ void k_export_hits(PrimaryQueue q, int n, igb200_hit* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 h = q.hit[i];
    igb200_hit o; o.ent_id = q.ent[i]; o.prim_id = __float_as_int(h.w); o.t = h.x; o.u = h.y; o.v = h.z;
    out[i] = o;
}

}  // namespace igb
