// C ABI (include/igb200.h) + global kernels + host-side wavefront loop of the B200 render device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is loaded at run time (igb200_comm_*)
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: named ranges on the host timeline for nsys / ncu --nvtx (the reference's own section timers, Statistics.h:25-55)

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cerrno>
#include <thread>
#include <sys/ipc.h>
#include <sys/shm.h>
#include <deque>
#include <string>
#include <vector>

#include "../../include/igb200.h"
#include "bvh8.h"
#include "bvh_build.h"
#include "wavefront.cuh"

using namespace igb;

// ================================================================================================ kernels
namespace {

constexpr int WF_BLOCK = 256;       // threads per CTA of the persistent kernels
// CTAs per SM the register budget is compiled for: 3 -> 80 registers / thread, 2 -> 128. Both variants are built; the
// "min_blocks" option picks one (DESIGN.md "Occupancy").
// "vote" picks the scheduling of the traversal visit kinds (traverse.cuh turn / turn_vote); all variants are built.
using WaveKernel = void (*)(const WaveParams);
using TraceKernel = void (*)(const DevScene, const PrimaryQueue, int, const ShadowQueue, int, float*, int*, unsigned long long*, int, int, int, int, int);
// "full": the shade code with every feature (shade.cuh shade_record<true>) or only what the BASELINE configurations use.
static WaveKernel wave_kernel(int min_blocks, int vote, bool full, int where = 0) {
    if (vote && where == 1) {   // scene staged as a whole (as turn_trace_kernel)
        if (full) return min_blocks >= 3 ? k_wavefront<WF_BLOCK, 3, 2, true, 1> : k_wavefront<WF_BLOCK, 2, 2, true, 1>;
        return min_blocks >= 3 ? k_wavefront<WF_BLOCK, 3, 2, false, 1> : k_wavefront<WF_BLOCK, 2, 2, false, 1>;
    }
    if (full) {
        if (vote) return min_blocks >= 3 ? k_wavefront<WF_BLOCK, 3, 2, true> : k_wavefront<WF_BLOCK, 2, 2, true>;
        return min_blocks >= 3 ? k_wavefront<WF_BLOCK, 3, 0, true> : k_wavefront<WF_BLOCK, 2, 0, true>;
    }
    if (vote) return min_blocks >= 3 ? k_wavefront<WF_BLOCK, 3, 2, false> : k_wavefront<WF_BLOCK, 2, 2, false>;
    return min_blocks >= 3 ? k_wavefront<WF_BLOCK, 3, 0, false> : k_wavefront<WF_BLOCK, 2, 0, false>;
}
static WaveKernel turn_shade_kernel(int blocks, bool full) {
    if (full) return blocks >= 4 ? k_turn_shade<WF_BLOCK, 4, true> : blocks == 3 ? k_turn_shade<WF_BLOCK, 3, true> : k_turn_shade<WF_BLOCK, 2, true>;
    return blocks >= 4 ? k_turn_shade<WF_BLOCK, 4, false> : blocks == 3 ? k_turn_shade<WF_BLOCK, 3, false> : k_turn_shade<WF_BLOCK, 2, false>;
}
// where: 1 the whole scene is staged in shared memory (shared-memory loads without range checks: -6 % trace time on diamond_scene
// and cbox), 0 decided per index (traverse.cuh node_ptr). Specialised for the default scheduling (vote 2) only. An "everything in
// global memory" variant (2) was measured too: +1 to +2.5 % (more spills), so unstaged scenes use the generic kernel.
// The merged-tree kernel of small scenes with its ray records staged through shared memory (wavefront.cuh phase_trace_staged): CTAs of
// 384 threads x 2 per SM or 768 x 1 (option "flat_block"), 256 = the unstaged kernel
static WaveKernel flat_staged_kernel(int block) { return block == 768 ? k_turn_trace<768, 1, 2, 1, true, true> : k_turn_trace<384, 2, 2, 1, true, true>; }
static WaveKernel turn_trace_kernel(int blocks, int vote, int where, bool flat = false) {
    if (flat && where == 2) return blocks >= 3 ? k_turn_trace<WF_BLOCK, 3, 2, 2, true> : k_turn_trace<WF_BLOCK, 2, 2, 2, true>;   // merged tree read from global memory (large scenes)
    if (flat) return blocks >= 3 ? k_turn_trace<WF_BLOCK, 3, 2, 1, true> : k_turn_trace<WF_BLOCK, 2, 2, 1, true>;   // merged tree (small scenes, staged)
    if (vote && where == 1) return blocks >= 3 ? k_turn_trace<WF_BLOCK, 3, 2, 1> : k_turn_trace<WF_BLOCK, 2, 2, 1>;
    if (vote && where == 2) return blocks >= 3 ? k_turn_trace<WF_BLOCK, 3, 2, 2> : k_turn_trace<WF_BLOCK, 2, 2, 2>;   // nothing staged: nodes through 256-bit global loads
    if (vote) return blocks >= 3 ? k_turn_trace<WF_BLOCK, 3, 2, 0> : k_turn_trace<WF_BLOCK, 2, 2, 0>;
    return blocks >= 3 ? k_turn_trace<WF_BLOCK, 3, 0, 0> : k_turn_trace<WF_BLOCK, 2, 0, 0>;
}
static TraceKernel trace_kernel(int min_blocks, int vote, int where = 0, bool flat = false) {
    if (flat) return min_blocks >= 3 ? k_trace<WF_BLOCK, 3, 2, 1, true> : k_trace<WF_BLOCK, 2, 2, 1, true>;
    if (vote && where == 1 && min_blocks < 4) return min_blocks >= 3 ? k_trace<WF_BLOCK, 3, 2, 1> : k_trace<WF_BLOCK, 2, 2, 1>;   // as turn_trace_kernel
    if (min_blocks >= 4) return vote ? k_trace<WF_BLOCK, 4, 2> : k_trace<WF_BLOCK, 4, 0>;   // 64 registers: stand-alone trace phase only
    if (vote) return min_blocks >= 3 ? k_trace<WF_BLOCK, 3, 2> : k_trace<WF_BLOCK, 2, 2>;
    return min_blocks >= 3 ? k_trace<WF_BLOCK, 3, 0> : k_trace<WF_BLOCK, 2, 0>;
}

// ---- the exchange step: pixels of one rank's tiles <-> a contiguous buffer [local tile][tile x tile][rgb]
// pack: this rank's k-th tile (tile_table[k] = row-major tile index) out of its accumulation buffer
__global__ void k_pack_tiles(const float* __restrict__ fb, float* __restrict__ out, const int* __restrict__ tile_table, int n_tiles, int tile, int tiles_x, int width, int height) {
    const long long n = (long long)n_tiles * tile * tile;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / (tile * tile)), r = (int)(i - (long long)k * tile * tile);
        const int t = tile_table[k];
        const int x = (t % tiles_x) * tile + r % tile, y = (t / tiles_x) * tile + r / tile;
        float3 v = make_float3(0, 0, 0);
        if (x < width && y < height) { const float* p = fb + ((size_t)y * width + x) * 3; v = make_float3(p[0], p[1], p[2]); }
        out[i * 3] = v.x; out[i * 3 + 1] = v.y; out[i * 3 + 2] = v.z;
    }
}
// unpack (rank 0): every pixel of the frame from the region of the rank that owns its tile; local_index[t] = position of tile t in its owner's list
__global__ void k_unpack_tiles(float* __restrict__ frame, const float* __restrict__ in, const int* __restrict__ local_index, const long long* __restrict__ rank_offset,
                               int world, int tile, int tiles_x, int width, int height) {
    const long long n = (long long)width * height;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % width), y = (int)(p / width);
        const int tx = x / tile, ty = y / tile, t = ty * tiles_x + tx;
        const int r = (tx + ty) % world;
        const float* src = in + rank_offset[r] + ((long long)local_index[t] * tile * tile + (long long)(y % tile) * tile + x % tile) * 3;
        frame[p * 3] = src[0]; frame[p * 3 + 1] = src[1]; frame[p * 3 + 2] = src[2];
    }
}

// Frame streaming: folds the finished iteration's slot into the accumulated frame, clears the slot for its next user and leaves a
// snapshot of the frame for the copy engine (the accumulated frame itself is folded into again while that copy runs)
__global__ void k_publish(float* __restrict__ acc, float* __restrict__ slot, float* __restrict__ snap, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float r = acc[i] + slot[i];
        acc[i] = r; slot[i] = 0.0f;
        if (snap) snap[i] = r;
    }
}

// Frame streaming over several ranks with the frames in SHARED host memory (igb200_frame_stream_share): every rank writes the pixels of its own
// tiles of the snapshot straight into the host frame (zero-copy stores into the mapped segment, each GPU over its own PCIe link), then raises
// its flag for that frame. Rank 0's host sees a frame complete when every rank's flag carries the frame's sequence number.
// VEC: 16-byte stores (tile x 3 and width x 3 floats are multiples of four, so every tile row starts on a 16-byte boundary): a warp store is 512
// contiguous bytes on the link instead of 128.
template <bool VEC>
__global__ void k_tiles_to_host(const float* __restrict__ snap, float* __restrict__ host_frame, const int* __restrict__ tile_table, int n_tiles, int tile, int tiles_x, int width, int height) {
    constexpr int V = VEC ? 4 : 1;
    const int row_words = tile * 3 / V;                               // one tile row = tile pixels x rgb, contiguous in the frame
    const long long n = (long long)n_tiles * tile * row_words;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / ((long long)tile * row_words));
        const int r = (int)(i - (long long)k * tile * row_words);
        const int ty = r / row_words, w = (r - ty * row_words) * V;
        const int t = tile_table ? tile_table[k] : k;
        const int x0 = (t % tiles_x) * tile, y = (t / tiles_x) * tile + ty;
        if (y >= height || x0 * 3 + w >= width * 3) continue;
        const size_t o = ((size_t)y * width + x0) * 3 + w;
        if (VEC) *reinterpret_cast<float4*>(host_frame + o) = *reinterpret_cast<const float4*>(snap + o);
        else host_frame[o] = snap[o];
    }
}
__global__ void k_raise_flag(volatile unsigned int* flag, unsigned int seq) {
    __threadfence_system();   // the frame's words (written by the kernel before this one on the same stream) before the flag
    *flag = seq;
    __threadfence_system();
}

// Deterministic accumulation: folds the per-sample slots of one iteration into the frame in sample order and clears them. One thread per
// word of the frame; the additions of a pixel happen in one thread in a fixed order, so the frame is a pure function of the inputs.
__global__ void k_resolve(float* __restrict__ frame, float* __restrict__ slots, long long n_pixels, int spi) {
    const long long n = n_pixels * 3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / 3; const int ch = (int)(i - p * 3);
        float acc = frame[i];
        for (int s = 0; s < spi; ++s) { float* w = slots + (p * spi + s) * 3 + ch; acc += *w; *w = 0.0f; }
        frame[i] = acc;
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
// NVTX range over a host-side step of the pipeline (no-ops unless a profiler is attached)
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } NvtxRange(const NvtxRange&) = delete; NvtxRange& operator=(const NvtxRange&) = delete; };
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
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device -> host copies of finished frames, overlapping the next iteration's kernels
    cudaEvent_t ev_frame = nullptr, ev_copied = nullptr;
    bool has_scene = false;
    bool scene_full = false;       // the scene uses features only shade_record<true> has (conductors, sphere / spot lights, non-uniform selectors)
    igb200_scene_desc desc{};      // scalar members only are kept
    DevScene dev{};
    DevBuf<float4> nodes, tris, ent_leaf, ent_shade, blob, materials;
    DevBuf<int> tri_prim;
    DevBuf<float4> flat_nodes;         // merged tree of small scenes (build_flat_tree)
    int flat_option = 1;               // option "flat": 0 = always walk the two-level tree, 1 = the merged tree when it fits shared memory, 2 = also when it does not (read from global memory)
    bool flat_on = false;              // the split-turn trace kernel and the trace hooks walk the merged tree
    bool flat_global = false;          // ... the split-turn trace kernel only, and the tree stays in global memory (large scenes, option "flat" = 2)
    int64_t flat_max_bytes = 1ll << 30;   // option "flat_max_mb": a merged tree larger than this is not built (instances multiply the nodes)
    size_t smem_flat = 0;
    // option "flat_block": 384 | 768 = the merged-tree trace kernel with its ray records staged through shared memory by TMA (CTAs of that size),
    // 256 = records read from global memory. Measured (profiles/r5b_staged_rays.txt): the staged kernel is SLOWER (diamond_scene trace 3.56 ->
    // 4.12 / 3.84 ms per step, cbox 2.44 -> 2.79) although it removes the long-scoreboard stalls of the refill (a quarter of the stall samples):
    // the other warps were covering them, and a swap every 32 rays costs more issue slots than the stalls did. Kept as an option, off.
    int flat_block_option = WF_BLOCK;
    int flat_block = WF_BLOCK;         // what runs: WF_BLOCK = the unstaged kernel
    size_t smem_flat_staged = 0;
    DevBuf<int4> shape_info;
    DevBuf<float> inf_lights, fin_lights, selector_data, textures, aux_data;
    DevBuf<int4> images;
    DevBuf<uint32_t> image_data;
    // framebuffer
    int width = 0, height = 0;
    DevBuf<float> fb;
    float* host_fb = nullptr; size_t host_fb_n = 0;
    // standard AOVs of the reference's infobuffer wrapper (technique/internal/infobuffer.art): [0] Normals, [1] Albedo; option "std_aovs"
    bool std_aovs = false;
    DevBuf<float> aov[2];
    float* host_aov[2] = {nullptr, nullptr};
    // queues
    size_t capacity = 0, want_capacity = (size_t)1 << 25;   // upper bound of records per queue (84 B each, two queues + 48 B shadow)
    QueueMem qa, qb;
    DevBuf<float4> sq_org, sq_dir, sq_col;
    DevBuf<int> bin_order;             // material binning: SHADE_BINS index lists of `capacity` entries
    int bin_materials = -1;            // option: 1 on, 0 off, -1 = on when the scene has a heavy material class (textures, maps, rough conductors)
    bool scene_heavy = false;
    DevBuf<Control> control;
    Control* host_control = nullptr;   // pinned
    // persistent-kernel configuration
    int blocks_per_sm = 0;             // resident CTAs per SM of k_wavefront with the current shared-memory size
    int stage_nodes = 0, stage_tris = 0, stage_ent = 0;
    size_t smem_bytes = 0;
    int64_t stage_budget = 40 * 1024;  // bytes of shared memory per CTA for the staged scene copy
    int turn_where = 0;                // ... of the split-turn trace kernel: as stage_where, or 2 = nothing staged (option "wide_loads")
    int wide_loads = 0;                // measured slower (+9 % trace on synthetic_room, profiles/r6_trace_experiments.txt): off
    int stage_where = 0;               // 1: the whole scene is staged in shared memory, 0: not (selects the k_turn_trace variant)
    int specialise_where = 1;          // option: 0 = always use the generic (per index) trace kernel
    int carveout = -1;                 // option: preferred shared-memory carve-out of the trace kernels in percent (-1: the driver's choice)
    int stage_partial = 0;             // 1: stage the prefix that fits even if the scene does not fit as a whole
    int refill = 24, min_blocks = 2, vote = 2;
    int drain_turns = -1;               // split turns at the start of a drain (-1: from the number of paths that may be waiting)
    int shade_sync = 1;                 // k_turn_shade keeps the warps of a CTA in step with block barriers (wavefront.cuh phase_shade_cta): 0 off, 1 on
    int wave_skip = 1;                  // a step ends with its split turns when nothing forces the persistent kernel to run (wavefront.cuh k_wavefront)
    int split_turns = -1;               // leading turns of an iteration run as separate shade / trace launches (0: all in the persistent kernel)
    int turn_trace_blocks = 3;         // CTAs per SM the trace kernel of a split turn is compiled for
    int turn_shade_blocks = 3;         // ... and the shade + generate kernel
    int grid_turn_shade = 0, grid_turn_trace = 0, grid_turn_trace_flat = 0;
    int trace_blocks = 0;              // stand-alone trace hooks: CTAs per SM the kernel is compiled for (0: as min_blocks)
    // deferred tail: a launch ends once at most defer_permille/1000 of the iteration's camera rays are still alive as paths;
    // they are carried into the next launch (render) or finished by a drain launch before anything is observed
    int defer_permille = 1000;
    int64_t wide_rays_per_group = 2;   // trace phases with at most this many rays per group of 8 lanes use the wide walk (0: never)
    bool pending = false;              // launches issued since the last synchronisation with the device
    // fused iterations: consecutive render() calls are collected and generated by ONE launch when a single iteration is too
    // small to fill the GPU (a 1/8 share of the frame at 8 GPUs); flushed by anything that observes or changes state
    std::vector<igb200_settings> queued;
    int fuse = 0;                      // iterations per launch (0: chosen so that a launch generates ~8 M camera rays, at most 8)
    bool maybe_carry = false;          // the last launch may have left paths behind
    long long last_defer = 0;          // ... at most this many
    igb200_settings carry_settings{};  // settings the carried paths were generated with
    int carry_rank = 0, carry_world = 1, carry_tile = 0;
    RenderParams last_rp{}; DevScene last_sc{};
    // partition
    int rank = 0, world = 1, tile = 32;
    DevBuf<int> tile_table;            // the rank's tiles (row-major tile indices), rebuilt when size / tile / rank / world change
    long long n_local_tiles = 0, tile_key[5] = {0, 0, 0, 0, 0};
    // stats
    uint64_t launches = 0;             // kernels launched by igb200_render (and its drains) since the last reset
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // "profile_kernels": every launch of igb200_render is bracketed by CUDA events on the stream (kind: 0 k_wavefront, 1 k_turn_trace,
    // 2 k_turn_shade, 3 k_turn_end); elapsed times are summed at the next synchronisation
    bool profile = false;
    struct Timed { cudaEvent_t a, b; int kind; };
    std::vector<Timed> timed;
    double prof_ms[4] = {0, 0, 0, 0}; uint64_t prof_n[4] = {0, 0, 0, 0};
    DevBuf<igb200_ray> list_rays;
    // BVH construction: which builder makes a shape's tree, and the on-disk cache (bvh_build.cu)
    int gpu_bvh = -1;                          // option "gpu_bvh": 1 = every shape of more than 4 faces is built on the GPU, 0 = never, -1 = from gpu_bvh_min_faces faces on
    int64_t gpu_bvh_min_faces = 1 << 20;       // option
    int64_t bvh_cache_min_faces = 500000;      // option: meshes with more faces go through the cache (the reference: MinFaceCountForCache, TriMeshProvider.cpp:328)
    std::string cache_dir;                     // igb200_set_cache_dir; empty = no cache
    int64_t build_info[6] = {0, 0, 0, 0, 0, 0};   // last igb200_set_scene: shapes built on the host / on the GPU / loaded from the cache / stored, microseconds spent, nodes
    // deterministic accumulation (option "deterministic"): per-sample slots for the colour buffer and the two standard AOVs
    bool deterministic = false;
    DevBuf<float> det_slots, det_aov[2];
    // frame streaming (igb200_frame_stream_*): per-iteration framebuffer slots + in-order publication of finished frames
    bool fs_on = false;
    int fs_slots = 0;                               // ring size (power of two)
    DevBuf<float> fs_ring, fs_snap[2];
    size_t fs_n4 = 0;                               // float4 per (padded) frame
    struct FsIter { int iter; long long shades_left; };
    std::deque<FsIter> fs_inflight;                 // iterations generated but not yet known to be finished, oldest first
    struct FsFrame { int iter; int host; cudaEvent_t copied; unsigned int seq; };
    std::deque<FsFrame> fs_ready;                   // published frames the caller has not taken yet, oldest first
    std::vector<float*> fs_host;                    // pinned frames
    std::vector<int> fs_host_free;
    std::vector<cudaEvent_t> fs_events;             // pool of "copied" events
    cudaEvent_t fs_ev_pub[2] = {nullptr, nullptr}, fs_ev_snapfree[2] = {nullptr, nullptr};
    bool fs_snap_used[2] = {false, false};
    long long fs_published = 0;
    int fs_last_taken_host = -1;
    // frames in shared host memory (igb200_frame_stream_share): a System V segment every rank attaches and pins
    int fs_share_key = 0;                           // 0: off
    int fs_shm_id = -1;
    unsigned char* fs_shm = nullptr;                // [header: consumed (64 B), flags[frame][64] u32] [frames]
    unsigned char* fs_shm_dev = nullptr;            // the same bytes as the GPU sees them
    size_t fs_shm_bytes = 0, fs_shm_frame_bytes = 0, fs_shm_header = 0;
    int fs_shm_frames = 0;                          // host frames in the segment
    bool fs_shm_retry = false;
    bool fs_taken_shared = false;                   // the caller holds a frame of the segment (released at the next call)
    // multi-GPU exchange (igb200_comm_*)
    ncclComm_t comm = nullptr;
    DevBuf<float> comm_send, comm_recv, comm_frame;
    DevBuf<int> comm_local_index;
    DevBuf<long long> comm_rank_offset;
    float* comm_host = nullptr; size_t comm_host_n = 0;
    long long comm_key[4] = {0, 0, 0, 0};
    std::vector<long long> comm_counts;   // floats each rank sends
};

static int ensure_queues(igb200_ctx* c, size_t need) {
    size_t cap = std::min(c->want_capacity, std::max<size_t>(need, 1024));
    cap = (cap + 1023) / 1024 * 1024;
    if (cap <= c->capacity) return 0;
    CU(c->qa.alloc(cap)); CU(c->qb.alloc(cap));
    CU(c->sq_org.alloc(cap)); CU(c->sq_dir.alloc(cap)); CU(c->sq_col.alloc(cap));
    c->bin_order.release();
    c->capacity = cap;
    return 0;
}

// Decides how much of the scene is staged into shared memory and how many CTAs of the persistent kernels fit an SM.
static int configure_kernels(igb200_ctx* c) {
    const DevScene& s = c->dev;
    int64_t left = c->stage_budget;
    // All or nothing: a partial copy costs more than it saves -- the shared memory it takes is L1 cache the rest of the scene then
    // misses (1920x1080x4: synthetic_room 10.34 -> 9.84 ms, primitives 2.98 -> 2.81 ms with nothing staged). "stage_partial" = 1
    // brings the greedy prefix back.
    if (!c->stage_partial && (int64_t)s.n_ent * STAGED_LEAF_BYTES + (int64_t)s.n_nodes * STAGED_NODE_BYTES + (int64_t)s.n_tris * 48 > left) left = 0;
    // entity leaves first (every ray reads them), then nodes (top of every tree first in memory order), then triangles
    c->stage_ent = (int)std::min<int64_t>(s.n_ent, left / STAGED_LEAF_BYTES); left -= (int64_t)c->stage_ent * STAGED_LEAF_BYTES;
    c->stage_nodes = (int)std::min<int64_t>(s.n_nodes, left / STAGED_NODE_BYTES); left -= (int64_t)c->stage_nodes * STAGED_NODE_BYTES;
    c->stage_tris = (int)std::min<int64_t>(s.n_tris, left / 48);
    c->stage_where = (c->stage_ent == s.n_ent && c->stage_nodes == s.n_nodes && c->stage_tris == s.n_tris) ? 1 : 0;
    if (!c->specialise_where) c->stage_where = 0;
    // the split-turn trace kernel also exists for "nothing staged" (large scenes): nodes come through 256-bit global loads (traverse.cuh node_step_global8)
    c->turn_where = (c->wide_loads && c->vote && c->stage_ent == 0 && c->stage_nodes == 0 && c->stage_tris == 0) ? 2 : c->stage_where;
    c->smem_bytes = (size_t)SMEM_STACK * WF_BLOCK * sizeof(uint2) + (size_t)c->stage_ent * STAGED_LEAF_BYTES + (size_t)c->stage_nodes * STAGED_NODE_BYTES + (size_t)c->stage_tris * 48;
    // the merged tree (small scenes): walked by the split-turn trace kernel and the trace hooks when the two-level scene is staged as a whole
    c->flat_on = c->flat_option != 0 && s.n_flat_nodes > 0 && c->stage_where == 1 && c->vote != 0 &&
                 (int64_t)s.n_flat_nodes * STAGED_NODE_BYTES + (int64_t)s.n_tris * 48 + (int64_t)s.n_ent * STAGED_LEAF_BYTES <= c->stage_budget;
    c->flat_global = !c->flat_on && c->flat_option >= 2 && s.n_flat_nodes > 0 && c->vote != 0 && c->stage_nodes == 0 && c->stage_tris == 0 && c->stage_ent == 0;
    if (c->flat_global) {
        const size_t smem = (size_t)SMEM_STACK * WF_BLOCK * sizeof(uint2);
        CU(cudaFuncSetAttribute((const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, 2, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nbg = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbg, (const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, 2, true), WF_BLOCK, smem));
        if (nbg < 1) c->flat_global = false; else c->grid_turn_trace_flat = nbg * c->n_sm;
    }
    c->smem_flat = (size_t)SMEM_STACK * WF_BLOCK * sizeof(uint2) + (size_t)s.n_ent * STAGED_LEAF_BYTES + (size_t)s.n_flat_nodes * STAGED_NODE_BYTES + (size_t)s.n_tris * 48;
    for (int full = 0; full < 2; ++full) CU(cudaFuncSetAttribute((const void*)wave_kernel(c->min_blocks, c->vote, full != 0, c->stage_where), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    CU(cudaFuncSetAttribute((const void*)trace_kernel(c->min_blocks, c->vote), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    int nb = 0;
    {   // both shade variants must fit the cooperative grid
        int nb1 = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)wave_kernel(c->min_blocks, c->vote, false, c->stage_where), WF_BLOCK, c->smem_bytes));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, (const void*)wave_kernel(c->min_blocks, c->vote, true, c->stage_where), WF_BLOCK, c->smem_bytes));
        nb = std::min(nb, nb1);
    }
    if (nb < 1) return fail(-2, "k_wavefront does not fit an SM with %zu bytes of shared memory", c->smem_bytes);
    c->blocks_per_sm = nb;
    // split turn kernels
    CU(cudaFuncSetAttribute((const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, c->turn_where), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    if (c->flat_on) {
        CU(cudaFuncSetAttribute((const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, 1, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_flat));
        int nbf = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbf, (const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, 1, true), WF_BLOCK, c->smem_flat));
        if (nbf < c->turn_trace_blocks) c->flat_on = false;   // would cost occupancy: keep the two-level walk
        else c->grid_turn_trace_flat = nbf * c->n_sm;
        c->flat_block = WF_BLOCK;
        if (c->flat_on && c->flat_block_option != WF_BLOCK) {   // ray records staged through shared memory, if the larger CTA fits as often as it must
            const int blk = c->flat_block_option, want = blk == 768 ? 1 : 2;
            const size_t scene = ((size_t)s.n_ent * STAGED_LEAF_BYTES + (size_t)s.n_flat_nodes * STAGED_NODE_BYTES + (size_t)s.n_tris * 48 + 127) & ~(size_t)127;
            const size_t smem = (size_t)SMEM_STACK * blk * sizeof(uint2) + scene + (size_t)(blk / 32) * STAGE_WARP_BYTES;
            int nbs = 0;
            if (cudaFuncSetAttribute((const void*)flat_staged_kernel(blk), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbs, (const void*)flat_staged_kernel(blk), blk, smem) == cudaSuccess && nbs >= want) {
                c->flat_block = blk; c->smem_flat_staged = smem; c->grid_turn_trace_flat = want * c->n_sm;
            } else cudaGetLastError();   // does not fit: the unstaged kernel stays
        }
    }
    if (c->carveout >= 0) {
        CU(cudaFuncSetAttribute((const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, c->turn_where), cudaFuncAttributePreferredSharedMemoryCarveout, c->carveout));
        for (int full = 0; full < 2; ++full) CU(cudaFuncSetAttribute((const void*)wave_kernel(c->min_blocks, c->vote, full != 0, c->stage_where), cudaFuncAttributePreferredSharedMemoryCarveout, c->carveout));
    }
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)turn_trace_kernel(c->turn_trace_blocks, c->vote, c->turn_where), WF_BLOCK, c->smem_bytes));
    if (nb < 1) return fail(-2, "k_turn_trace does not fit an SM with %zu bytes of shared memory", c->smem_bytes);
    c->grid_turn_trace = nb * c->n_sm;
    {
        int nb1 = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)turn_shade_kernel(c->turn_shade_blocks, false), WF_BLOCK, 0));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, (const void*)turn_shade_kernel(c->turn_shade_blocks, true), WF_BLOCK, 0));
        nb = std::min(nb, nb1);
    }
    if (nb < 1) return fail(-2, "k_turn_shade does not fit an SM");
    c->grid_turn_shade = nb * c->n_sm;
    return 0;
}

static WaveParams make_params(igb200_ctx* c, const RenderParams& rp, const DevScene& sc, long long total, const igb200_ray* d_rays, int defer) {
    WaveParams P;
    P.sc = sc; P.rp = rp;
    P.rp.ring_mask = 0; P.rp.ring_stride = 0;
    if (c->fs_on && !d_rays) { P.rp.ring_mask = c->fs_slots - 1; P.rp.ring_stride = (long long)c->fb.n; }
    P.sc.full = (c->scene_full || rp.aov_normals != nullptr) ? 1 : 0;
    P.q[0] = c->qa.view(); P.q[1] = c->qb.view();
    P.sq = ShadowQueue{c->sq_org.p, c->sq_dir.p, c->sq_col.p};
    P.fb = P.rp.det ? c->det_slots.p : P.rp.ring_stride ? c->fs_ring.p : c->fb.p; P.ctl = c->control.p;
    P.total = total; P.capacity = (int)c->capacity; P.list_rays = d_rays;
    P.stage_nodes = c->stage_nodes; P.stage_tris = c->stage_tris; P.stage_ent = c->stage_ent;
    P.refill = c->refill; P.defer = defer;
    P.shade_sync = c->shade_sync; P.wave_skip = c->wave_skip;
    P.order = c->bin_order.p;
    P.wide_limit = (int)std::min<int64_t>(c->wide_rays_per_group * c->blocks_per_sm * c->n_sm * (WF_BLOCK / 8), (int64_t)1 << 30);
    P.stage_flat = c->flat_on ? sc.n_flat_nodes : 0;
    return P;
}


static int prof_begin(igb200_ctx* c, int kind) {
    if (!c->profile) return 0;
    igb200_ctx::Timed t; t.kind = kind;
    CU(cudaEventCreate(&t.a)); CU(cudaEventCreate(&t.b));
    CU(cudaEventRecord(t.a, c->stream));
    c->timed.push_back(t);
    return 0;
}
static int prof_end(igb200_ctx* c) {
    if (!c->profile) return 0;
    CU(cudaEventRecord(c->timed.back().b, c->stream));
    return 0;
}
static int prof_collect(igb200_ctx* c) {   // after a stream synchronisation
    for (const igb200_ctx::Timed& t : c->timed) {
        float ms = 0; CU(cudaEventElapsedTime(&ms, t.a, t.b));
        c->prof_ms[t.kind] += ms; c->prof_n[t.kind] += 1;
        cudaEventDestroy(t.a); cudaEventDestroy(t.b);
    }
    c->timed.clear();
    return 0;
}

// One cooperative launch of the persistent kernel on the context's stream; asynchronous.
static int launch_wave(igb200_ctx* c, const RenderParams& rp, const DevScene& sc, long long total, const igb200_ray* d_rays, int defer) {
    WaveParams P = make_params(c, rp, sc, total, d_rays, defer);
    CU(cudaMemsetAsync(c->control.p, 0, CONTROL_SCRATCH, c->stream));
    void* args[] = {&P};
    { const int r = prof_begin(c, 0); if (r) return r; }
    CU(cudaLaunchCooperativeKernel((const void*)wave_kernel(c->min_blocks, c->vote, P.sc.full != 0, c->stage_where), dim3((unsigned)(c->blocks_per_sm * c->n_sm)), dim3(WF_BLOCK), args, c->smem_bytes, c->stream));
    { const int r = prof_end(c); if (r) return r; }
    c->launches += 1;
    c->pending = true;
    return 0;
}

// Finishes the paths earlier launches left behind (deferred tail): one more launch without camera rays that runs every
// path to its end. Everything that observes results (framebuffer, statistics) or changes what carried records refer to
// (scene, size, partition, spi) calls this first.
static int ensure_aovs(igb200_ctx* c) {
    if (!c->std_aovs || !c->fb.n) return 0;
    for (int k = 0; k < 2; ++k) {
        if (c->aov[k].n == c->fb.n) continue;
        CU(c->aov[k].alloc(c->fb.n));
        CU(cudaMemset(c->aov[k].p, 0, c->fb.n * sizeof(float)));
        if (c->host_aov[k]) { cudaFreeHost(c->host_aov[k]); c->host_aov[k] = nullptr; }
        CU(cudaMallocHost(&c->host_aov[k], c->fb.n * sizeof(float)));
    }
    return 0;
}

// `turns` wavefront turns as ordinary launches (shade + generate, trace, hand-over), see wavefront.cuh
static int launch_split_turns(igb200_ctx* c, const WaveParams& P, int turns) {
    if (turns <= 0) return 0;
    CU(cudaMemsetAsync(c->control.p, 0, CONTROL_SCRATCH, c->stream));
    for (int t = 0; t < turns; ++t) {
        { const int r = prof_begin(c, 2); if (r) return r; }
        turn_shade_kernel(c->turn_shade_blocks, P.sc.full != 0)<<<c->grid_turn_shade, WF_BLOCK, 0, c->stream>>>(P);
        { const int r = prof_end(c); if (r) return r; }
        { const int r = prof_begin(c, 1); if (r) return r; }
        if (c->flat_global) {   // the merged tree in global memory: the kernel's node array IS the merged tree, nothing is staged
            WaveParams G = P; G.sc.nodes = G.sc.flat_nodes; G.stage_flat = 0; G.stage_nodes = 0; G.stage_tris = 0; G.stage_ent = 0;
            turn_trace_kernel(c->turn_trace_blocks, c->vote, 2, true)<<<c->grid_turn_trace_flat, WF_BLOCK, (size_t)SMEM_STACK * WF_BLOCK * sizeof(uint2), c->stream>>>(G);
        }
        else if (c->flat_on && c->flat_block != WF_BLOCK) flat_staged_kernel(c->flat_block)<<<c->grid_turn_trace_flat, c->flat_block, c->smem_flat_staged, c->stream>>>(P);
        else if (c->flat_on) turn_trace_kernel(c->turn_trace_blocks, c->vote, 1, true)<<<c->grid_turn_trace_flat, WF_BLOCK, c->smem_flat, c->stream>>>(P);
        else turn_trace_kernel(c->turn_trace_blocks, c->vote, c->turn_where)<<<c->grid_turn_trace, WF_BLOCK, c->smem_bytes, c->stream>>>(P);
        { const int r = prof_end(c); if (r) return r; }
        { const int r = prof_begin(c, 3); if (r) return r; }
        k_turn_end<<<1, 1, 0, c->stream>>>(P);
        { const int r = prof_end(c); if (r) return r; }
        c->launches += 3;
    }
    CU(cudaGetLastError());
    c->pending = true;
    return 0;
}

// This rank's tiles (row-major tile indices) on the device: tile (tx, ty) belongs to rank (tx + ty) mod world
static int ensure_tile_table(igb200_ctx* c, int W, int H) {
    const int tiles_x = (W + c->tile - 1) / c->tile, tiles_y = (H + c->tile - 1) / c->tile;
    const long long tiles_total = (long long)tiles_x * tiles_y;
    const long long key[5] = {W, H, c->tile, c->rank, c->world};
    if (std::memcmp(key, c->tile_key, sizeof(key)) == 0) return 0;
    std::vector<int> table;
    for (long long t = 0; t < tiles_total; ++t) if ((int)((t % tiles_x + t / tiles_x) % c->world) == c->rank) table.push_back((int)t);
    CU(cudaStreamSynchronize(c->stream));
    CU(c->tile_table.upload(table));
    c->n_local_tiles = (long long)table.size();
    std::memcpy(c->tile_key, key, sizeof(key));
    return 0;
}

// iterations per launch: enough to generate ~8 M camera rays, at most 8 (option "fuse" overrides)
static int fuse_factor(const igb200_ctx* c, const igb200_settings* st) {
    if (c->deterministic) return 1;   // every iteration is resolved on its own
    if (c->fuse > 0) return c->fuse;
    const long long cam_rays = std::max<long long>((long long)st->width * st->height * st->spi / std::max(c->world, 1), 1);
    return (int)std::min<long long>(8, std::max<long long>(1, ((long long)1 << 23) / cam_rays));
}
extern "C" { static int launch_iterations(igb200_ctx* c, const igb200_settings* st, int n_iter, const igb200_ray* rays, size_t n_rays); }
static int flush_queued(igb200_ctx* c) {
    if (c->queued.empty()) return 0;
    const igb200_settings first = c->queued.front();
    const int n = (int)c->queued.size();
    c->queued.clear();   // before the launch: it may call drain / sync_control itself
    return launch_iterations(c, &first, n, nullptr, 0);
}

extern "C" { static int fs_publish(igb200_ctx* c, bool all); }
static int drain(igb200_ctx* c) {
    NvtxRange nvtx_("igb200: drain (deferred tail)");
    { const int r = flush_queued(c); if (r) return r; }
    if (c->maybe_carry) {
        CU(cudaSetDevice(c->device));
        // up to `last_defer` paths may be waiting: their first turns are still big enough for the split kernels
        const int turns = c->drain_turns >= 0 ? c->drain_turns : c->split_turns >= 0 ? c->split_turns : (int)std::min<long long>(8, c->last_defer >> 19);   // measured: 9.90 -> 9.51 ms per drained step with 4 -> 12
        { const int r = launch_split_turns(c, make_params(c, c->last_rp, c->last_sc, 0, nullptr, 0), turns); if (r) return r; }
        const int r = launch_wave(c, c->last_rp, c->last_sc, 0, nullptr, 0);
        if (r) return r;
        c->maybe_carry = false;
    }
    // frame streaming: with every path finished, every iteration still in flight is complete -> fold them all into the frame
    if (c->fs_on && !c->fs_inflight.empty()) { CU(cudaSetDevice(c->device)); const int r = fs_publish(c, true); if (r) return r; }
    return 0;
}

// drain + wait for the device + fetch the control block (statistics, logs)
static int sync_control(igb200_ctx* c) {
    { const int r = drain(c); if (r) return r; }
    if (c->pending) {
        CU(cudaSetDevice(c->device));
        CU(cudaMemcpyAsync(c->host_control, c->control.p, sizeof(Control), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        c->pending = false;
        { const int r = prof_collect(c); if (r) return r; }
    }
    return 0;
}

extern "C" {

const char* igb200_last_error(void) { return g_error.c_str(); }

int igb200_version(int* major, int* minor) {
    if (major) *major = IGB200_VERSION_MAJOR;
    if (minor) *minor = IGB200_VERSION_MINOR;
    return 0;
}

int igb200_device_count(int* count) {
    if (!count) return fail(-1, "igb200_device_count: count is null");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    int usable = 0;
    for (int d = 0; d < n; ++d) { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) usable = d + 1; }   // a prefix: contexts are addressed by CUDA ordinal
    *count = usable;
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
    if (!prop.cooperativeLaunch) return fail(-3, "igb200_create: device %d does not support cooperative launches", cuda_device);
    igb200_ctx* c = new igb200_ctx();
    c->device = cuda_device;
    c->n_sm = prop.multiProcessorCount;
    const int rc = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CU(c->control.alloc(1));
        CU(cudaMemset(c->control.p, 0, sizeof(Control)));
        CU(cudaMemset(&c->control.p->turn_t0, 0xFF, 2 * sizeof(unsigned long long)));
        CU(cudaMallocHost(&c->host_control, sizeof(Control)));
        std::memset(c->host_control, 0, sizeof(Control));
        CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1));
        CU(cudaEventCreateWithFlags(&c->ev_frame, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
        return 0;
    }();
    if (rc) { igb200_destroy(c); return rc; }   // nothing leaks on the error paths
    *out = c;
    return 0;
}

int igb200_destroy(igb200_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->fs_on) igb200_frame_stream_end(c);
    for (cudaEvent_t e : c->fs_events) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) { if (c->fs_ev_pub[k]) cudaEventDestroy(c->fs_ev_pub[k]); if (c->fs_ev_snapfree[k]) cudaEventDestroy(c->fs_ev_snapfree[k]); }
    igb200_comm_destroy(c);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_frame) cudaEventDestroy(c->ev_frame);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    if (c->host_fb) cudaFreeHost(c->host_fb);
    for (int k = 0; k < 2; ++k) if (c->host_aov[k]) cudaFreeHost(c->host_aov[k]);
    if (c->host_control) cudaFreeHost(c->host_control);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int igb200_set_option(igb200_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return fail(-1, "igb200_set_option: null argument");
    { const int r = sync_control(c); if (r) return r; }
    if (!strcmp(name, "capacity")) { if (value < 1024) return fail(-1, "capacity must be >= 1024"); c->want_capacity = (size_t)value; c->capacity = 0; return 0; }
    if (!strcmp(name, "vote")) { if (value < 0 || value > 2) return fail(-1, "vote must be 0 or 2"); c->vote = value ? 2 : 0; if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); } return 0; }
    if (!strcmp(name, "wide_rays_per_group")) { if (value < 0) return fail(-1, "wide_rays_per_group must be >= 0"); c->wide_rays_per_group = value; return 0; }
    if (!strcmp(name, "defer_permille")) { if (value < 0 || value > 8000) return fail(-1, "defer_permille must be in [0, 8000]"); c->defer_permille = (int)value; return 0; }
    if (!strcmp(name, "profile_kernels")) { c->profile = value != 0; return 0; }
    if (!strcmp(name, "wide_loads")) { if (value < 0 || value > 1) return fail(-1, "wide_loads must be 0 or 1"); { const int r = sync_control(c); if (r) return r; } c->wide_loads = (int)value; return c->has_scene ? configure_kernels(c) : 0; }
    if (!strcmp(name, "drain_turns")) { if (value < -1 || value > 64) return fail(-1, "drain_turns must be in [-1, 64]"); c->drain_turns = (int)value; return 0; }
    if (!strcmp(name, "shade_sync")) { if (value < 0 || value > 1) return fail(-1, "shade_sync must be 0 or 1"); c->shade_sync = (int)value; return 0; }
    if (!strcmp(name, "wave_skip")) { if (value < 0 || value > 1) return fail(-1, "wave_skip must be 0 or 1"); { const int r = drain(c); if (r) return r; } c->wave_skip = (int)value; return 0; }
    if (!strcmp(name, "refill")) { if (value < 1 || value > 32) return fail(-1, "refill must be in [1, 32]"); c->refill = (int)value; return 0; }
    if (!strcmp(name, "std_aovs")) { c->std_aovs = value != 0; CU(cudaSetDevice(c->device)); return ensure_aovs(c); }
    if (!strcmp(name, "flat")) {   // 0: never walk the merged tree. Takes effect at the next igb200_set_scene (the tree is built there) or at once when switching off
        if (value < 0 || value > 2) return fail(-1, "flat must be 0 (never), 1 (when the merged tree fits shared memory) or 2 (also from global memory)");
        c->flat_option = (int)value;
        if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); }
        return 0;
    }
    if (!strcmp(name, "gpu_bvh")) { if (value < -1 || value > 1) return fail(-1, "gpu_bvh must be -1 (from gpu_bvh_min_faces faces on), 0 (never) or 1 (every shape of more than 4 faces)"); c->gpu_bvh = (int)value; return 0; }
    if (!strcmp(name, "gpu_bvh_min_faces")) { if (value < 5) return fail(-1, "gpu_bvh_min_faces must be >= 5"); c->gpu_bvh_min_faces = value; return 0; }
    if (!strcmp(name, "bvh_cache_min_faces")) { if (value < 0) return fail(-1, "bvh_cache_min_faces must be >= 0"); c->bvh_cache_min_faces = value; return 0; }
    if (!strcmp(name, "flat_max_mb")) { if (value < 0 || value > (1 << 16)) return fail(-1, "flat_max_mb must be in [0, 65536]"); c->flat_max_bytes = value << 20; return 0; }
    if (!strcmp(name, "flat_block")) {
        if (value != 256 && value != 384 && value != 768) return fail(-1, "flat_block must be 256 (ray records read from global memory), 384 or 768 (staged through shared memory)");
        c->flat_block_option = (int)value;
        if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); }
        return 0;
    }
    if (!strcmp(name, "bin_materials")) { if (value < -1 || value > 2) return fail(-1, "bin_materials must be -1 (automatic), 0, 1 or 2"); c->bin_materials = (int)value; return 0; }
    if (!strcmp(name, "deterministic")) {
        // Same inputs => bit-identical frame, whatever the spi (src/tests/integrator/test_reproducibility.py:5-11; the reference's CPU device adds
        // without atomics, driver/accumulator.art:4-21). Costs one slot per sample (12 B x W x H x spi) and the deferred tail / fused iterations.
        if (c->fs_on && value) return fail(-1, "deterministic accumulation and frame streaming exclude each other (end the frame stream first)");
        c->deterministic = value != 0;
        if (!c->deterministic) { c->det_slots.release(); c->det_aov[0].release(); c->det_aov[1].release(); }
        return 0;
    }
    if (!strcmp(name, "fuse")) { if (value < 0 || value > 64) return fail(-1, "fuse must be in [0, 64]"); c->fuse = (int)value; return 0; }
    if (!strcmp(name, "split_turns")) { if (value < -1 || value > 64) return fail(-1, "split_turns must be in [-1, 64] (-1: chosen from the number of camera rays)"); c->split_turns = (int)value; return 0; }
    if (!strcmp(name, "turn_shade_blocks")) {
        if (value < 2 || value > 4) return fail(-1, "turn_shade_blocks must be 2, 3 or 4");
        c->turn_shade_blocks = (int)value;
        if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); }
        return 0;
    }
    if (!strcmp(name, "turn_trace_blocks")) {
        if (value != 2 && value != 3) return fail(-1, "turn_trace_blocks must be 2 or 3");
        c->turn_trace_blocks = (int)value;
        if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); }
        return 0;
    }
    if (!strcmp(name, "trace_blocks")) { if (value != 0 && (value < 2 || value > 4)) return fail(-1, "trace_blocks must be 0, 2, 3 or 4"); c->trace_blocks = (int)value; return 0; }
    if (!strcmp(name, "min_blocks")) {
        if (value != 2 && value != 3) return fail(-1, "min_blocks must be 2 or 3");
        c->min_blocks = (int)value;
        if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); }
        return 0;
    }
    if (!strcmp(name, "specialise_where")) { c->specialise_where = value ? 1 : 0; if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); } return 0; }
    if (!strcmp(name, "carveout")) { if (value < -1 || value > 100) return fail(-1, "carveout must be -1 or 0..100"); c->carveout = (int)value; if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); } return 0; }
    if (!strcmp(name, "stage_partial")) { c->stage_partial = value ? 1 : 0; if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); } return 0; }
    if (!strcmp(name, "stage_budget")) {
        if (value < 0 || value > 160 * 1024) return fail(-1, "stage_budget must be in [0, 163840] bytes");
        c->stage_budget = value;
        if (c->has_scene) { CU(cudaSetDevice(c->device)); return configure_kernels(c); }
        return 0;
    }
    return fail(-1, "igb200_set_option: unknown option '%s'", name);
}

int igb200_set_partition(igb200_ctx* c, int rank, int world, int tile_size) {
    if (!c) return fail(-1, "null context");
    if (world < 1 || rank < 0 || rank >= world || tile_size < 1) return fail(-1, "igb200_set_partition: invalid rank %d / world %d / tile %d", rank, world, tile_size);
    if (c->fs_on) return fail(-1, "igb200_set_partition: end the frame stream first (igb200_frame_stream_end)");
    { const int r = sync_control(c); if (r) return r; }
    c->rank = rank; c->world = world; c->tile = tile_size;
    return 0;
}

int igb200_set_scene(igb200_ctx* c, const igb200_scene_desc* d) {
    NvtxRange nvtx_("igb200_set_scene (BVH build + upload)");
    if (!c || !d) return fail(-1, "igb200_set_scene: null argument");
    CU(cudaSetDevice(c->device));
    { const int r = sync_control(c); if (r) return r; }
    if (d->technique.max_depth > 254) return fail(-4, "igb200_set_scene: max_depth %d > 254 is not supported (the depth travels in 8 bits of the ray record)", d->technique.max_depth);
    {   // light selector and the buffer it reads
        const int sel = d->technique.light_selector, n = d->n_finite;
        if (sel < IGB200_SELECTOR_UNIFORM || sel > IGB200_SELECTOR_HIERARCHY) return fail(-4, "igb200_set_scene: unsupported light selector %d", sel);
        if (sel != IGB200_SELECTOR_UNIFORM) {
            if (n < 1) return fail(-1, "igb200_set_scene: light selector %d needs finite lights (the reference generates the uniform selector then, LoaderLight.cpp:428-437)", sel);
            const long long need = sel == IGB200_SELECTOR_CDF ? n : (n == 1 ? 0 : (long long)(n + 3) / 4 * 4 + 8LL * (2 * n - 1));
            if (!d->selector_data || d->n_selector_data < need) return fail(-1, "igb200_set_scene: light selector %d over %d finite lights needs %lld words of selector_data, got %d", sel, n, need, d->n_selector_data);
            if (sel == IGB200_SELECTOR_HIERARCHY && n > 1) {   // every child index must stay inside the buffer; depth <= 32 (the codes are 32 bits)
                const int n_nodes = (d->n_selector_data - (n + 3) / 4 * 4) / 8;
                const float* e = d->selector_data + (n + 3) / 4 * 4;
                for (int k = 0; k < n_nodes; ++k) {
                    int32_t idx; std::memcpy(&idx, e + 8 * k + 7, 4);
                    if (idx >= n || (idx < 0 && (-idx - 1 + 1 >= n_nodes || -idx - 1 <= k))) return fail(-1, "igb200_set_scene: light hierarchy node %d refers to %d (%d nodes, %d lights)", k, idx, n_nodes, n);
                }
            }
        }
    }
    if (d->n_leaves != d->n_entities) return fail(-1, "igb200_set_scene: %d leaves for %d entities (one EntityLeaf1 per entity expected)", d->n_leaves, d->n_entities);
    if (d->shape_data_bytes % 16) return fail(-1, "igb200_set_scene: shapes dyn-table data must be a multiple of 16 bytes");
    if (d->n_textures < 0 || d->n_images < 0 || d->n_aux_data < 0 || (d->n_textures && !d->textures) || (d->n_images && !d->images) || (d->n_aux_data && !d->aux_data))
        return fail(-1, "igb200_set_scene: bad texture / image / aux_data tables");
    for (int i = 0; i < d->n_images; ++i) {
        const igb200_image& im = d->images[i];
        if (im.format < IGB200_IMAGE_RGBA8 || im.format > IGB200_IMAGE_RGBA32F || im.width < 1 || im.height < 1 || !im.pixels || (long long)im.width * im.height > (1ll << 28))
            return fail(-1, "igb200_set_scene: image %d is invalid (format %d, %d x %d)", i, im.format, im.width, im.height);
    }
    for (int t = 0; t < d->n_textures; ++t) {
        const igb200_texture& tx = d->textures[t];
        if (tx.type != IGB200_TEX_CHECKERBOARD && tx.type != IGB200_TEX_IMAGE) return fail(-4, "igb200_set_scene: texture %d has unsupported type %d", t, tx.type);
        if (tx.type == IGB200_TEX_IMAGE && (tx.image < 0 || tx.image >= d->n_images || tx.filter < 0 || tx.filter > 2 || tx.border_u < 0 || tx.border_u > 2 || tx.border_v < 0 || tx.border_v > 2))
            return fail(-1, "igb200_set_scene: texture %d refers to a bad image / filter / border", t);
    }
    auto tex_ok = [&](int id) { return id < d->n_textures; };   // negative: no texture
    for (int m = 0; m < d->n_materials; ++m) {
        const igb200_material& mt = d->materials[m];
        if (mt.bsdf < IGB200_BSDF_DIFFUSE || mt.bsdf > IGB200_BSDF_CONDUCTOR) return fail(-4, "igb200_set_scene: material %d has unsupported bsdf %d", m, mt.bsdf);
        if (!tex_ok(mt.tex[0]) || !tex_ok(mt.tex[1])) return fail(-1, "igb200_set_scene: material %d refers to a texture that does not exist", m);
        if (mt.map_kind < IGB200_MAP_NONE || mt.map_kind > IGB200_MAP_NORMAL || (mt.map_kind != IGB200_MAP_NONE && (mt.map_tex < 0 || !tex_ok(mt.map_tex))))
            return fail(-1, "igb200_set_scene: material %d has a bad bump / normal map", m);
        if (mt.distribution != IGB200_MICROFACET_DELTA && mt.distribution != IGB200_MICROFACET_VNDF_GGX) return fail(-4, "igb200_set_scene: material %d has unsupported microfacet distribution %d", m, mt.distribution);
    }
    for (int l = 0; l < d->n_infinite; ++l) {
        const igb200_light& li = d->infinite_lights[l];
        const int t = li.type;
        if (t != IGB200_LIGHT_ENV_CONST && t != IGB200_LIGHT_SUN && t != IGB200_LIGHT_DIRECTIONAL && t != IGB200_LIGHT_ENV_TEXTURED && t != IGB200_LIGHT_ENV_TEX) return fail(-4, "igb200_set_scene: infinite light %d has unsupported type %d", l, t);
        if (t == IGB200_LIGHT_ENV_TEXTURED || t == IGB200_LIGHT_ENV_TEX) {
            int32_t w[4]; std::memcpy(w, li.p + 12, sizeof(w));
            if (w[0] < 0 || w[0] >= d->n_textures) return fail(-1, "igb200_set_scene: environment light %d refers to texture %d (%d textures)", l, w[0], d->n_textures);
            if (t == IGB200_LIGHT_ENV_TEXTURED && (w[1] < 0 || w[2] < 1 || w[3] < 1 || (long long)w[1] + (long long)w[3] * (w[2] + 1) > d->n_aux_data))
                return fail(-1, "igb200_set_scene: environment light %d: its cdf (%d x %d at word %d) does not fit aux_data (%d words)", l, w[2], w[3], w[1], d->n_aux_data);
        }
    }
    for (int l = 0; l < d->n_finite; ++l) {
        const int t = d->finite_lights[l].type;
        if (t != IGB200_LIGHT_POINT && t != IGB200_LIGHT_PLANE_AREA && t != IGB200_LIGHT_SHAPE_AREA && t != IGB200_LIGHT_SPHERE_AREA && t != IGB200_LIGHT_SPOT) return fail(-4, "igb200_set_scene: finite light %d has unsupported type %d", l, t);
    }

    // ---- per-shape geometry: triangles in BVH leaf order + BVH8 (replaces the reference's pre-baked trimesh_primbvh table)
    for (int64_t& v : c->build_info) v = 0;
    std::vector<Node8> nodes;
    std::vector<float4> tris;
    std::vector<int> tri_prim;
    std::vector<int4> shape_info(2 * (size_t)d->n_shapes);
    std::vector<int> shape_root(d->n_shapes, 0);
    std::vector<int> shape_node_base(d->n_shapes, 0), shape_node_count(d->n_shapes, 0);   // the shape's tree inside `nodes`, before the reordering below
    std::vector<Box3> ent_boxes(d->n_entities);
    for (int i = 0; i < d->n_entities; ++i) {
        const igb200_entity_leaf& lf = d->leaves[i];
        for (int k = 0; k < 3; ++k) { ent_boxes[i].lo[k] = lf.min[k]; ent_boxes[i].hi[k] = lf.max[k]; }
    }
    // top-level tree first so that its root is node 1
    Bvh8 top = build_bvh8(ent_boxes, 1);
    int max_shape_depth = 0;
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
        if (lk.offset + 48 > d->shape_data_bytes) return fail(-1, "igb200_set_scene: shape %d: trimesh header beyond the end of the shapes table", s);
        const int nf = h[0], nv = h[1], nn = h[2], nt = h[3];
        // shapes/trimesh.art:77-96: header 16 B + box 32 B, then 16 B per vertex, normal and face, 8 B per texture coordinate
        if (nf < 0 || nv < 0 || nn < 0 || nt < 0 || (uint64_t)lk.offset + 48 + 16ull * ((uint64_t)nv + (uint64_t)nn + (uint64_t)nf) + 8ull * (uint64_t)nt > d->shape_data_bytes)
            return fail(-1, "igb200_set_scene: shape %d: trimesh header (%d faces, %d vertices, %d normals, %d texcoords) does not fit the shapes table", s, nf, nv, nn, nt);
        // normals and texture coordinates are looked up with the vertex indices (shapes/trimesh.art:20-36)
        if (nf > 0 && (nn < nv || nt < nv)) return fail(-1, "igb200_set_scene: shape %d has fewer normals (%d) or texture coordinates (%d) than vertices (%d)", s, nn, nt, nv);
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
        Bvh8 bvh;
        {   // the shape's tree: from the cache, from the GPU builder or from the host builder (bvh_build.cu)
            const auto t0 = std::chrono::steady_clock::now();
            const bool on_gpu = nf > 4 && (c->gpu_bvh == 1 || (c->gpu_bvh < 0 && nf >= c->gpu_bvh_min_faces));
            const int builder = on_gpu ? BVH_BUILDER_GPU_LBVH : BVH_BUILDER_HOST_SAH;
            const bool cached = !c->cache_dir.empty() && nf > c->bvh_cache_min_faces;
            const uint64_t hash = cached ? bvh_cache_hash(boxes, 4, builder) : 0;
            bool have = cached && bvh_cache_load(c->cache_dir, hash, (size_t)nf, builder, bvh);
            const bool loaded = have;
            if (loaded) c->build_info[2]++;
            if (!have && on_gpu) {
                std::string err;
                have = build_bvh8_gpu(boxes, bvh, c->stream, err);
                if (have) c->build_info[1]++;
                else { cudaGetLastError(); fprintf(stderr, "[igb200] shape %d: %s -- building on the host\n", s, err.c_str()); bvh = Bvh8(); }
            }
            if (!have) { bvh = build_bvh8(boxes, 4); c->build_info[0]++; }
            // a tree the GPU builder could not make is stored under the GPU builder's key all the same: the key names the request, and any valid tree answers it
            if (cached && !loaded && bvh_cache_store(c->cache_dir, hash, builder, bvh)) c->build_info[3]++;
            c->build_info[4] += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
            c->build_info[5] += (int64_t)bvh.nodes.size();
        }
        max_shape_depth = std::max(max_shape_depth, nf > 4 ? bvh.max_depth : 0);
        const int node_base = (int)nodes.size(), tri_base = (int)(tris.size() / 3);
        shape_root[s] = node_base + 1;
        if (nf == 0) {   // empty mesh: a root without children
            Node8 e;
            for (int k = 0; k < 6; ++k) for (int c8 = 0; c8 < 8; ++c8) e.bounds[k][c8] = (k & 1) ? -FLT_MAX : FLT_MAX;
            for (int c8 = 0; c8 < 8; ++c8) { e.child[c8] = 0; e.pad[c8] = 0; }
            bvh.nodes.push_back(e);
        }
        if (nf >= 1 && nf <= 4) {
            // a shape of at most four triangles is one leaf: the entity leaf points straight at it, no inner node
            bvh.nodes.clear();
            for (int t = 0; t < nf; ++t) bvh.order[t] = t;
            shape_root[s] = -(((tri_base << 2) | (nf - 1)) + 1);
        }
        shape_node_base[s] = node_base; shape_node_count[s] = (int)bvh.nodes.size();
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
            tris.push_back(make_float4(p0[0], p0[1], p0[2], n[0]));
            tris.push_back(make_float4(e1[0], e1[1], e1[2], n[1]));
            tris.push_back(make_float4(e2[0], e2[1], e2[2], n[2]));
            tri_prim.push_back(t);
        }
    }
    {   // Traversal-stack bound: a visit of an inner node leaves at most 7 entries behind, entering an instance one sentinel. Beyond
        // the stack the walks would drop entries silently (missed hits, light leaks), so such a scene is refused instead.
        const int need = 7 * (top.max_depth + max_shape_depth) + 2;
        const int have = std::min(STACK_SIZE, WIDE_STACK);
        if (need > have) return fail(-4, "igb200_set_scene: BVH too deep for the traversal stack (top level %d + shape level %d levels need %d entries, %d available)", top.max_depth, max_shape_depth, need, have);
    }
    // ---- merged tree for small scenes (traverse.cuh "merged-tree walk"): the top-level tree whose leaves are replaced by the instances'
    // own trees, each instance's node boxes refitted in WORLD space from its transformed triangles. Triangles stay per shape (local space).
    std::vector<Node8> flat;
    {
        bool ok = c->flat_option != 0 && d->n_entities >= 1 && d->n_entities <= 510 && tri_prim.size() < ((size_t)1 << 20);
        for (int s = 0; s < d->n_shapes && ok; ++s) ok = d->shape_lookups[s].type_id == IGB200_SHAPE_TRIMESH;
        size_t flat_nodes_n = top.nodes.size();
        for (int i = 0; i < d->n_entities && ok; ++i) flat_nodes_n += shape_root[d->leaves[i].shape_id] > 0 ? shape_node_count[d->leaves[i].shape_id] : 0;
        ok = ok && ((int64_t)flat_nodes_n * STAGED_NODE_BYTES + (int64_t)tri_prim.size() * 48 + (int64_t)d->n_entities * STAGED_LEAF_BYTES <= c->stage_budget ||   // all of it staged in shared memory ...
                    (c->flat_option >= 2 && (int64_t)flat_nodes_n * 256 <= c->flat_max_bytes));                                                 // ... or walked in global memory
        if (ok) {
            flat.assign(top.nodes.begin(), top.nodes.end());
            const float inf = std::numeric_limits<float>::max();
            for (size_t tn = 0; tn < top.nodes.size(); ++tn) {
                for (int k = 0; k < 8; ++k) {
                    const int code = flat[tn].child[k];
                    if (code >= 0) continue;                              // inner child of the top-level tree (indices are already those of `flat`)
                    const int slot = (-code - 1) >> 2;                    // top-level leaves hold one entity each
                    const igb200_entity_leaf& lf = d->leaves[top.order[slot]];
                    const int sh = lf.shape_id, ent = lf.entity_id & 0x7FFFFFFF;
                    const float* G = d->entities + 36 * (size_t)ent + 12;   // local -> world, column major 3x4 (LoaderEntity.cpp:150-162)
                    const uint8_t* sp = d->shape_data + d->shape_lookups[sh].offset;
                    const int32_t* hdr = reinterpret_cast<const int32_t*>(sp);
                    const float* verts = reinterpret_cast<const float*>(sp) + 12;
                    const int32_t* inds = reinterpret_cast<const int32_t*>(verts + 4 * (size_t)hdr[1] + 4 * (size_t)hdr[2]);
                    // world-space box of the triangle in primitive slot `ts`, padded well beyond the rounding of transform + triangle test
                    auto tri_box = [&](int ts) {
                        Box3 b = Box3::empty();
                        const int t = tri_prim[ts];
                        for (int v = 0; v < 3; ++v) {
                            const float* p = verts + 4 * (size_t)inds[4 * t + v];
                            float w[3];
                            for (int a = 0; a < 3; ++a) w[a] = (float)((double)G[a] * p[0] + (double)G[3 + a] * p[1] + (double)G[6 + a] * p[2] + (double)G[9 + a]);
                            b.extend(w);
                        }
                        for (int a = 0; a < 3; ++a) {
                            const float pad = 3.0517578125e-05f * std::max(std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])), b.hi[a] - b.lo[a]) + 1e-30f;   // 2^-15 relative
                            b.lo[a] -= pad; b.hi[a] += pad;
                        }
                        return b;
                    };
                    auto leaf_code = [&](int shape_leaf_code, Box3& box) {   // the shape's leaf code -> the merged tree's, + the leaf's world box
                        const int r = -shape_leaf_code - 1, first = r >> 2, cnt = (r & 3) + 1;
                        box = Box3::empty();
                        for (int j = 0; j < cnt; ++j) box.extend(tri_box(first + j));
                        return -((((slot << 20) | first) << 2 | (cnt - 1)) + 1);
                    };
                    auto set_box = [&](Node8& n, int lane, const Box3& b) {
                        n.bounds[0][lane] = b.lo[0]; n.bounds[1][lane] = b.hi[0]; n.bounds[2][lane] = b.lo[1]; n.bounds[3][lane] = b.hi[1]; n.bounds[4][lane] = b.lo[2]; n.bounds[5][lane] = b.hi[2];
                    };
                    Box3 root_box;
                    if (shape_root[sh] < 0) { flat[tn].child[k] = leaf_code(shape_root[sh], root_box); set_box(flat[tn], k, root_box); continue; }
                    // copy the shape's tree, children rebased, boxes refitted bottom-up (children follow their parents in the array)
                    const int base = (int)flat.size(), sb = shape_node_base[sh], sn = shape_node_count[sh];
                    for (int i = 0; i < sn; ++i) flat.push_back(nodes[sb + i]);
                    std::vector<Box3> node_box(sn, Box3::empty());
                    for (int i = sn - 1; i >= 0; --i) {
                        Node8& n = flat[base + i];
                        Box3 nb = Box3::empty();
                        for (int ch = 0; ch < 8; ++ch) {
                            Box3 cb;
                            if (n.child[ch] > 0) { const int ci = n.child[ch] - 1 - sb; cb = node_box[ci]; n.child[ch] = base + ci + 1; }
                            else if (n.child[ch] < 0) n.child[ch] = leaf_code(n.child[ch], cb);
                            else { for (int q = 0; q < 6; ++q) n.bounds[q][ch] = (q & 1) ? -inf : inf; continue; }
                            set_box(n, ch, cb); nb.extend(cb);
                        }
                        node_box[i] = nb;
                    }
                    flat[tn].child[k] = base + 1;
                    set_box(flat[tn], k, node_box[0]);
                }
            }
        }
    }
    // ---- node order: breadth-first across ALL trees (top-level root first, then level by level over every shape's tree), so that
    // the part of the node array that is staged in shared memory (configure_kernels) holds the top of every tree instead of
    // the whole first shape: with geometry that does not fit, every bottom-level walk then saves its first L2 round trips
    if (!nodes.empty()) {
        const int n = (int)nodes.size();
        std::vector<int> order; order.reserve(n);
        std::vector<char> seen(n, 0);
        std::vector<int> frontier, next;
        frontier.push_back(0);
        for (int s = 0; s < d->n_shapes; ++s) if (shape_root[s] > 0) frontier.push_back(shape_root[s] - 1);
        while (!frontier.empty()) {
            next.clear();
            for (const int i : frontier) {
                if (i < 0 || i >= n || seen[i]) continue;
                seen[i] = 1; order.push_back(i);
                for (int k = 0; k < 8; ++k) if (nodes[i].child[k] > 0) next.push_back(nodes[i].child[k] - 1);
            }
            frontier.swap(next);
        }
        for (int i = 0; i < n; ++i) if (!seen[i]) order.push_back(i);
        std::vector<int> perm(n);
        for (int k = 0; k < n; ++k) perm[order[k]] = k;
        std::vector<Node8> sorted(n);
        for (int i = 0; i < n; ++i) {
            Node8 nd = nodes[i];
            for (int k = 0; k < 8; ++k) if (nd.child[k] > 0) nd.child[k] = perm[nd.child[k] - 1] + 1;
            sorted[perm[i]] = nd;
        }
        nodes.swap(sorted);
        for (int s = 0; s < d->n_shapes; ++s) if (shape_root[s] > 0) shape_root[s] = perm[shape_root[s] - 1] + 1;
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
        const float* m = lf.local;  // column major 3x4 -> rows
        // bit 1: the local matrix is bit for bit the identity, so transforming the ray is the exact map x -> x + 0
        static const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
        const bool identity = std::memcmp(m, ident, sizeof(ident)) == 0;
        // bit 2: only a translation (linear part bit for bit the identity): the direction, hence its reciprocals, stay as they are
        const bool translation = !identity && std::memcmp(m, ident, 9 * sizeof(float)) == 0;
        L[1] = make_float4(lf.max[0], lf.max[1], lf.max[2], as_f((type == IGB200_SHAPE_SPHERE ? 1 : 0) | (identity ? 2 : 0) | (translation ? 4 : 0)));
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
        // shade bin of the entity's material (wavefront.cuh bin_append): 0 plain, 1 heavy (textured parameters, bump / normal map, rough conductor)
        const igb200_material& mt = d->materials[mat_id];
        const bool heavy = mt.tex[0] >= 0 || mt.tex[1] >= 0 || mt.map_kind != IGB200_MAP_NONE ||
                           (mt.bsdf == IGB200_BSDF_CONDUCTOR && mt.distribution == IGB200_MICROFACET_VNDF_GGX && mt.alpha_u > 1e-4f && mt.alpha_v > 1e-4f);
        // (experiment "bin_materials" = 2: when no material is heavy, split delta BSDFs -- no next-event estimation -- from the others)
        E[5] = make_float4(nm[2], nm[5], nm[8], as_f((heavy || (c->bin_materials == 2 && mt.bsdf != IGB200_BSDF_DIFFUSE)) ? 1 : 0));
    }
    std::vector<float4> flat_f4(flat.size() * 16);
    if (!flat.empty()) std::memcpy(flat_f4.data(), flat.data(), flat.size() * sizeof(Node8));
    std::vector<float4> node_f4(nodes.size() * 16);
    if (!nodes.empty()) std::memcpy(node_f4.data(), nodes.data(), nodes.size() * sizeof(Node8));
    std::vector<float4> blob(d->shape_data_bytes / 16);
    if (!blob.empty()) std::memcpy(blob.data(), d->shape_data, d->shape_data_bytes);
    std::vector<float4> mats(8 * (size_t)d->n_materials);
    static_assert(sizeof(igb200_material) == 128 && sizeof(igb200_texture) == 96, "descriptor layout");
    std::vector<float> texs(24 * (size_t)d->n_textures);
    if (d->n_textures) std::memcpy(texs.data(), d->textures, sizeof(igb200_texture) * (size_t)d->n_textures);
    std::vector<int4> img_table((size_t)d->n_images);
    std::vector<uint32_t> img_words;
    for (int i = 0; i < d->n_images; ++i) {
        const igb200_image& im = d->images[i];
        const size_t bytes = (size_t)im.width * im.height * (im.format == IGB200_IMAGE_RGBA8 ? 4 : im.format == IGB200_IMAGE_MONO8 ? 1 : 16);
        const size_t first = img_words.size();                      // a multiple of 4 words: float4 pixels stay 16-byte aligned
        img_words.resize(first + (bytes + 15) / 16 * 4, 0u);
        std::memcpy(img_words.data() + first, im.pixels, bytes);
        if (first > 0x7fffffffull) return fail(-1, "igb200_set_scene: more than 8 GB of image data");
        img_table[i] = make_int4(im.format, im.width, im.height, (int)first);
    }
    std::vector<float> aux;
    if (d->n_aux_data) aux.assign(d->aux_data, d->aux_data + d->n_aux_data);
    if (d->n_materials) std::memcpy(mats.data(), d->materials, sizeof(igb200_material) * (size_t)d->n_materials);
    std::vector<float> infl(32 * (size_t)d->n_infinite), finl(32 * (size_t)d->n_finite);
    if (d->n_infinite) std::memcpy(infl.data(), d->infinite_lights, sizeof(igb200_light) * (size_t)d->n_infinite);
    if (d->n_finite) std::memcpy(finl.data(), d->finite_lights, sizeof(igb200_light) * (size_t)d->n_finite);

    CU(cudaStreamSynchronize(c->stream));
    CU(c->flat_nodes.upload(flat_f4));
    CU(c->nodes.upload(node_f4)); CU(c->tris.upload(tris)); CU(c->tri_prim.upload(tri_prim)); CU(c->ent_leaf.upload(ent_leaf)); CU(c->ent_shade.upload(ent_shade));
    CU(c->blob.upload(blob)); CU(c->shape_info.upload(shape_info)); CU(c->materials.upload(mats));
    CU(c->inf_lights.upload(infl)); CU(c->fin_lights.upload(finl));
    CU(c->textures.upload(texs)); CU(c->images.upload(img_table)); CU(c->image_data.upload(img_words)); CU(c->aux_data.upload(aux));
    std::vector<float> seld;
    if (d->technique.light_selector != IGB200_SELECTOR_UNIFORM) seld.assign(d->selector_data, d->selector_data + d->n_selector_data);
    CU(c->selector_data.upload(seld));

    DevScene& s = c->dev;
    s.nodes = c->nodes.p; s.tris = c->tris.p; s.tri_prim = c->tri_prim.p; s.ent_leaf = c->ent_leaf.p; s.ent_shade = c->ent_shade.p; s.blob = c->blob.p;
    s.shape_info = c->shape_info.p; s.materials = c->materials.p; s.inf_lights = c->inf_lights.p; s.fin_lights = c->fin_lights.p;
    s.textures = c->textures.p; s.images = c->images.p; s.image_data = c->image_data.p; s.aux_data = c->aux_data.p;
    s.n_ent = d->n_entities; s.n_mat = d->n_materials; s.n_inf = d->n_infinite; s.n_fin = d->n_finite;
    s.n_nodes = (int)nodes.size(); s.n_tris = (int)tri_prim.size();
    s.flat_nodes = c->flat_nodes.p; s.n_flat_nodes = (int)flat.size();
    {   // bbox_radius(scene_bbox) * 1.01: light/env.art:76, core/bbox.art:24 (fma dot as on the device)
        const float dx = d->bbox_max[0] - d->bbox_min[0], dy = d->bbox_max[1] - d->bbox_min[1], dz = d->bbox_max[2] - d->bbox_min[2];
        s.scene_radius = std::sqrt(std::fmaf(dx, dx, std::fmaf(dy, dy, dz * dz))) / 2 * 1.01f;
    }
    s.selector = d->technique.light_selector; s.selector_data = c->selector_data.p;
    c->scene_full = s.selector != IGB200_SELECTOR_UNIFORM;
    for (int m = 0; m < d->n_materials; ++m) {
        const igb200_material& mt = d->materials[m];
        c->scene_full |= mt.bsdf == IGB200_BSDF_CONDUCTOR || mt.tex[0] >= 0 || mt.tex[1] >= 0 || mt.map_kind != IGB200_MAP_NONE;
    }
    for (int l = 0; l < d->n_finite; ++l) c->scene_full |= d->finite_lights[l].type == IGB200_LIGHT_SPHERE_AREA || d->finite_lights[l].type == IGB200_LIGHT_SPOT;
    for (int l = 0; l < d->n_infinite; ++l) c->scene_full |= d->infinite_lights[l].type != IGB200_LIGHT_ENV_CONST;
    s.full = c->scene_full ? 1 : 0;
    c->scene_heavy = false;
    for (int e = 0; e < d->n_entities; ++e) c->scene_heavy |= __builtin_bit_cast(int, ent_shade[6 * (size_t)e + 5].w) != 0;
    s.max_depth = d->technique.max_depth; s.min_depth = d->technique.min_depth; s.clamp_value = d->technique.clamp; s.nee = d->technique.nee;
    c->desc = *d;
    c->desc.entities = nullptr; c->desc.shape_lookups = nullptr; c->desc.shape_data = nullptr; c->desc.leaves = nullptr;
    c->desc.entity_per_material = nullptr; c->desc.materials = nullptr; c->desc.infinite_lights = nullptr; c->desc.finite_lights = nullptr;
    c->desc.selector_data = nullptr; c->desc.textures = nullptr; c->desc.images = nullptr; c->desc.aux_data = nullptr;
    c->has_scene = true;
    return configure_kernels(c);
}

int igb200_set_cache_dir(igb200_ctx* c, const char* dir) {
    if (!c) return fail(-1, "null context");
    c->cache_dir = dir ? dir : "";
    while (c->cache_dir.size() > 1 && c->cache_dir.back() == '/') c->cache_dir.pop_back();
    return 0;
}

int igb200_scene_build_info(igb200_ctx* c, int64_t out[6]) {
    if (!c || !out) return fail(-1, "igb200_scene_build_info: null argument");
    for (int k = 0; k < 6; ++k) out[k] = c->build_info[k];
    return 0;
}

int igb200_resize(igb200_ctx* c, int width, int height) {
    if (!c) return fail(-1, "null context");
    if (width < 1 || height < 1) return fail(-1, "igb200_resize: invalid size %dx%d", width, height);
    CU(cudaSetDevice(c->device));
    if (width == c->width && height == c->height) return 0;
    if (c->fs_on) return fail(-1, "igb200_resize: end the frame stream first (igb200_frame_stream_end)");
    { const int r = sync_control(c); if (r) return r; }
    c->width = width; c->height = height;
    std::memset(c->comm_key, 0, sizeof(c->comm_key));
    const size_t n = (size_t)width * height * 3;
    CU(c->fb.alloc(n));
    CU(cudaMemset(c->fb.p, 0, n * sizeof(float)));
    if (c->host_fb) { cudaFreeHost(c->host_fb); c->host_fb = nullptr; }
    CU(cudaMallocHost(&c->host_fb, n * sizeof(float)));
    c->host_fb_n = n;
    for (int k = 0; k < 2; ++k) {
        c->aov[k].release();
        if (c->host_aov[k]) { cudaFreeHost(c->host_aov[k]); c->host_aov[k] = nullptr; }
    }
    { const int r = ensure_aovs(c); if (r) return r; }
    return 0;
}

static bool is_color(const char* aov) { return !aov || !*aov || !strcmp(aov, "Color"); }
// 0: colour, 1: Normals, 2: Albedo (when "std_aovs" is on), -1: no such AOV
static int aov_index(const igb200_ctx* c, const char* aov) {
    if (is_color(aov)) return 0;
    if (c->std_aovs && !strcmp(aov, "Normals")) return 1;
    if (c->std_aovs && !strcmp(aov, "Albedo")) return 2;
    return -1;
}

int igb200_clear(igb200_ctx* c, const char* aov) {
    if (!c) return fail(-1, "null context");
    const int which = aov ? aov_index(c, aov) : 0;   // NULL: clearAllFramebuffer
    if (which < 0) return fail(-4, "igb200_clear: AOV '%s' does not exist", aov);
    CU(cudaSetDevice(c->device));
    // paths still in flight belong to the image that is being thrown away: finish them first, then clear
    { const int r = drain(c); if (r) return r; }
    if (c->fb.p && (!aov || which == 0)) CU(cudaMemsetAsync(c->fb.p, 0, c->fb.n * sizeof(float), c->stream));
    for (int k = 0; k < 2; ++k) if (c->aov[k].p && (!aov || which == k + 1)) CU(cudaMemsetAsync(c->aov[k].p, 0, c->aov[k].n * sizeof(float), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int igb200_framebuffer(igb200_ctx* c, const char* aov, float** host_ptr) {
    NvtxRange nvtx_("igb200_framebuffer (drain + D2H)");
    if (!c || !host_ptr) return fail(-1, "igb200_framebuffer: null argument");
    const int which = aov_index(c, aov);
    if (which < 0) return fail(-4, "igb200_framebuffer: AOV '%s' does not exist", aov);
    { const int r = flush_queued(c); if (r) return r; }   // a queued first render creates the framebuffer
    if (!c->fb.p) return fail(-1, "igb200_framebuffer: no framebuffer (call igb200_resize first)");
    CU(cudaSetDevice(c->device));
    { const int r = drain(c); if (r) return r; }
    float* dst = which == 0 ? c->host_fb : c->host_aov[which - 1];
    const float* src = which == 0 ? c->fb.p : c->aov[which - 1].p;
    if (!src || !dst) return fail(-1, "igb200_framebuffer: AOV '%s' is not allocated", aov);
    CU(cudaMemcpyAsync(dst, src, c->fb.n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *host_ptr = dst;
    return 0;
}

int igb200_framebuffer_device(igb200_ctx* c, const char* aov, float** device_ptr) {
    if (!c || !device_ptr) return fail(-1, "igb200_framebuffer_device: null argument");
    const int which = aov_index(c, aov);
    if (which < 0) return fail(-4, "igb200_framebuffer_device: AOV '%s' does not exist", aov);
    { const int r = flush_queued(c); if (r) return r; }
    if (!c->fb.p) return fail(-1, "igb200_framebuffer_device: no framebuffer");
    CU(cudaSetDevice(c->device));
    { const int r = drain(c); if (r) return r; }
    CU(cudaStreamSynchronize(c->stream));
    *device_ptr = which == 0 ? c->fb.p : c->aov[which - 1].p;
    return 0;
}

int igb200_upload_framebuffer(igb200_ctx* c, const char* aov, const float* host_rgb) {
    if (!c || !host_rgb) return fail(-1, "igb200_upload_framebuffer: null argument");
    const int which = aov_index(c, aov);
    if (which < 0) return fail(-4, "igb200_upload_framebuffer: AOV '%s' does not exist", aov);
    { const int r = flush_queued(c); if (r) return r; }
    if (!c->fb.p) return fail(-1, "igb200_upload_framebuffer: no framebuffer");
    CU(cudaSetDevice(c->device));
    { const int r = drain(c); if (r) return r; }
    CU(cudaMemcpyAsync(which == 0 ? c->fb.p : c->aov[which - 1].p, host_rgb, c->fb.n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int igb200_stream(igb200_ctx* c, void** cuda_stream) {
    if (!c || !cuda_stream) return fail(-1, "igb200_stream: null argument");
    // whatever the caller enqueues next is ordered after every render() so far INCLUDING its deferred tail: the drain launches are
    // asynchronous, so this does not wait for the device (a gather ordered after flush_queued() alone read an incomplete frame)
    CU(cudaSetDevice(c->device));
    { const int r = drain(c); if (r) return r; }
    *cuda_stream = (void*)c->stream;
    return 0;
}

int igb200_sync(igb200_ctx* c) {
    if (!c) return fail(-1, "null context");
    return sync_control(c);
}

int igb200_stats(igb200_ctx* c, uint64_t out[5], double* render_ms) {
    if (!c) return fail(-1, "null context");
    { const int r = sync_control(c); if (r) return r; }
    const Control& h = *c->host_control;
    if (out) { out[0] = h.stat[0]; out[1] = h.stat[1]; out[2] = h.stat[2]; out[3] = h.stat[3] + h.trace_splats; out[4] = c->launches; }
    if (render_ms) *render_ms = (double)h.kernel_ns * 1e-6;
    return 0;
}

int igb200_reset_stats(igb200_ctx* c) {
    if (!c) return fail(-1, "null context");
    { const int r = sync_control(c); if (r) return r; }
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->control.p, 0, sizeof(Control), c->stream));   // nothing is carried after a drain
    CU(cudaMemsetAsync(&c->control.p->turn_t0, 0xFF, 2 * sizeof(unsigned long long), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::memset(c->host_control, 0, sizeof(Control));
    c->launches = 0;
    for (int k = 0; k < 4; ++k) { c->prof_ms[k] = 0; c->prof_n[k] = 0; }
    return 0;
}

int igb200_launch_profile(igb200_ctx* c, double ms[4], uint64_t launches[4], uint64_t split_work[3]) {
    if (!c) return fail(-1, "null context");
    { const int r = sync_control(c); if (r) return r; }
    for (int k = 0; k < 4; ++k) { if (ms) ms[k] = c->prof_ms[k]; if (launches) launches[k] = c->prof_n[k]; }
    if (split_work) for (int k = 0; k < 3; ++k) split_work[k] = c->host_control->split[k];
    return 0;
}

int igb200_turn_log(igb200_ctx* c, uint32_t* items, uint32_t* trace_ns, uint32_t* shade_ns, int max_turns, int* n_turns) {
    if (!c || !items || !trace_ns || !shade_ns || !n_turns) return fail(-1, "igb200_turn_log: null argument");
    // no drain here: the log describes the last launch as it ran (a drain would overwrite it)
    { const int r = flush_queued(c); if (r) return r; }
    if (c->pending) {
        CU(cudaSetDevice(c->device));
        CU(cudaMemcpyAsync(c->host_control, c->control.p, sizeof(Control), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        c->pending = c->maybe_carry;
    }
    const Control& h = *c->host_control;
    const int n = (int)std::min<unsigned long long>(std::min<unsigned long long>(h.turns, TURN_LOG), (unsigned long long)std::max(max_turns, 0));
    for (int k = 0; k < n; ++k) { items[k] = h.turn_items[k]; trace_ns[k] = h.turn_trace_ns[k]; shade_ns[k] = h.turn_shade_ns[k]; }
    *n_turns = n;
    return 0;
}

int igb200_step_stats(igb200_ctx* c, uint64_t out[16]) {
    if (!c || !out) return fail(-1, "igb200_step_stats: null argument");
    { const int r = sync_control(c); if (r) return r; }
    for (int k = 0; k < 16; ++k) out[k] = c->host_control->step_stats[k / 8][k % 8];
    return 0;
}

int igb200_kernel_times(igb200_ctx* c, double out_ms[4], uint64_t out_launches[4]) {
    if (!c) return fail(-1, "null context");
    { const int r = sync_control(c); if (r) return r; }
    const Control& h = *c->host_control;
    const double ms[4] = {0, (double)h.phase_ns[0] * 1e-6, (double)h.phase_ns[1] * 1e-6, 0};
    const uint64_t n[4] = {0, h.phases[0], h.phases[1], 0};
    for (int k = 0; k < 4; ++k) { if (out_ms) out_ms[k] = ms[k]; if (out_launches) out_launches[k] = n[k]; }
    return 0;
}

static int launch_iterations(igb200_ctx* c, const igb200_settings* st, int n_iter, const igb200_ray* rays, size_t n_rays) {
    NvtxRange nvtx_("igb200_render: launch iterations");
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
    { const int r = ensure_aovs(c); if (r) return r; }
    rp.aov_normals = (c->std_aovs && !rays) ? c->aov[0].p : nullptr; rp.aov_albedo = (c->std_aovs && !rays) ? c->aov[1].p : nullptr;
    rp.det = c->deterministic ? 1 : 0;
    if (rp.det) {   // one slot per sample; allocated zeroed, k_resolve leaves them zeroed
        if (n_iter != 1) return fail(-1, "igb200_render: deterministic accumulation renders one iteration per launch");
        const size_t n_slots = (size_t)W * H * st->spi * 3;
        auto want = [&](DevBuf<float>& b) -> int { if (b.n == n_slots) return 0; { const int r = sync_control(c); if (r) return r; } CU(b.alloc(n_slots)); CU(cudaMemset(b.p, 0, n_slots * sizeof(float))); return 0; };
        { const int r = want(c->det_slots); if (r) return r; }
        if (rp.aov_normals) { for (int k = 0; k < 2; ++k) { const int r = want(c->det_aov[k]); if (r) return r; } rp.aov_normals = c->det_aov[0].p; rp.aov_albedo = c->det_aov[1].p; }
    }
    if (rays) { rp.tile_w = W; rp.tile_h = 1; rp.rank = 0; rp.world = 1; }
    else { rp.tile_w = c->tile; rp.tile_h = c->tile; rp.rank = c->rank; rp.world = c->world; }
    rp.tiles_x = (W + rp.tile_w - 1) / rp.tile_w;
    const int tiles_y = (H + rp.tile_h - 1) / rp.tile_h;
    const long long tiles_total = (long long)rp.tiles_x * tiles_y;
    // this rank's tiles: tile (tx, ty) belongs to rank (tx + ty) mod world -- diagonals, so that no rank is tied to a set of columns or
    // rows; the list is kept on the device (phase_generate maps its k-th local tile through it)
    long long local_tiles = tiles_total;
    rp.tile_table = nullptr;
    if (rp.world > 1) {
        { const int r = ensure_tile_table(c, W, H); if (r) return r; }
        local_tiles = c->n_local_tiles;
        rp.tile_table = c->tile_table.p;
    }
    rp.per_iter = local_tiles * rp.tile_w * rp.tile_h * rp.spi;             // padded ray domain of this rank, one iteration
    const long long total = rp.per_iter * n_iter;                          // ... of the launch (n_iter consecutive iterations)

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
    // Paths carried over from the previous launch are shaded with this launch's parameters: anything they depend on
    // must be unchanged, else they are finished first.
    const igb200_settings& cs = c->carry_settings;
    const bool compatible = !rays && cs.spi == st->spi && cs.width == W && cs.height == H && cs.frame == st->frame && cs.seed == st->seed &&
                            c->carry_rank == rp.rank && c->carry_world == rp.world && c->carry_tile == rp.tile_w;
    if (c->maybe_carry && !compatible) { const int r = drain(c); if (r) return r; }
    if (st->iter < 0 || st->iter + n_iter > (1 << 24)) return fail(-1, "igb200_render: iteration %d out of range [0, 2^24)", st->iter);
    const igb200_ray* d_rays = nullptr;
    if (rays) {
        { const int r = sync_control(c); if (r) return r; }   // the previous list may still be read
        CU(c->list_rays.alloc(n_rays));
        CU(cudaMemcpyAsync(c->list_rays.p, rays, n_rays * sizeof(igb200_ray), cudaMemcpyHostToDevice, c->stream));
        d_rays = c->list_rays.p;
    }
    // With a deferred tail the launch returns once at most `defer` paths are alive; they continue in the next launch or in
    // a drain launch. The queues are sized so that the carried records and all new camera rays fit the first turn.
    const long long cam_rays = (long long)W * H * st->spi / std::max(rp.world, 1) * n_iter;
    const long long want_defer = (rays || rp.det) ? 0 : cam_rays * c->defer_permille / 1000;   // deterministic: every path ends inside its own launch
    // queues sized for a full batch of fused iterations from the start (a reallocation costs tens of milliseconds)
    const int f_full = rays ? 1 : std::max(n_iter, fuse_factor(c, st));
    const size_t need = (size_t)std::max<long long>((total / n_iter + want_defer / n_iter) * f_full, 1);
    if (std::min(need, c->want_capacity) > c->capacity) {   // the queues are about to be reallocated
        { const int r = sync_control(c); if (r) return r; }
        { const int r = ensure_queues(c, need); if (r) return r; }
    }
    const int defer = (int)std::min<long long>(want_defer, (long long)c->capacity - 1024);
    {   // material binning (a9): on by default when shading cost differs between material classes
        const bool want_bins = !rays && (c->bin_materials >= 1 || (c->bin_materials < 0 && c->scene_heavy));
        if (want_bins && !c->bin_order.p) { { const int r = sync_control(c); if (r) return r; } CU(c->bin_order.alloc((size_t)SHADE_BINS * c->capacity)); }
        if (!want_bins && c->bin_order.p) { { const int r = sync_control(c); if (r) return r; } c->bin_order.release(); }
    }

    // The iteration (asynchronously: nothing comes back to the host): its first, big turns as split launches, then one
    // cooperative launch of the persistent kernel that generates whatever camera rays did not fit yet and runs until at
    // most `defer` paths are alive.
    // frame streaming: every iteration in flight owns a framebuffer slot; when the ring would overflow, everything is finished first
    if (c->fs_on && !rays && (int)c->fs_inflight.size() + n_iter > c->fs_slots) { const int r = drain(c); if (r) return r; }
    CU(cudaMemsetAsync(&c->control.p->next_cam, 0, sizeof(long long), c->stream));
    // split turns pay three launches each: worth it while a turn holds millions of rays (4 at 8 M camera rays, 2 at 1 M)
    const int split_turns = c->split_turns >= 0 ? c->split_turns : (int)std::min<long long>(4, std::max<long long>(1, cam_rays >> 19));
    if (!rays) { const int r = launch_split_turns(c, make_params(c, rp, sc, total, nullptr, defer), split_turns); if (r) return r; }
    { const int r = launch_wave(c, rp, sc, total, d_rays, defer); if (r) return r; }
    if (rp.det) {
        k_resolve<<<c->n_sm * 8, 256, 0, c->stream>>>(c->fb.p, c->det_slots.p, (long long)W * H, st->spi);
        if (rp.aov_normals && st->iter == 0) for (int k = 0; k < 2; ++k) k_resolve<<<c->n_sm * 8, 256, 0, c->stream>>>(c->aov[k].p, c->det_aov[k].p, (long long)W * H, st->spi);
        CU(cudaGetLastError());
        c->launches += 1;
    }
    c->maybe_carry = defer > 0; c->last_defer = defer;
    c->carry_settings = *st; c->carry_settings.width = W; c->carry_settings.height = H; c->carry_settings.iter = st->iter + n_iter - 1;
    c->carry_rank = rp.rank; c->carry_world = rp.world; c->carry_tile = rp.tile_w;
    c->last_rp = rp; c->last_sc = sc;
    if (c->fs_on && !rays) {
        // Which iterations are finished now? Every launch shades the WHOLE primary queue at least (split turns + 1) times, a path is
        // dead after max_depth shades (technique/pathtracer.art:68,177), and shadow rays never outlive the turn that made them: an
        // iteration generated by an earlier launch is complete once the launches after it add up to max_depth shade passes. Exact, and
        // known to the host without a read-back; the same on every rank.
        const long long shades = (long long)split_turns + (c->wave_skip ? 0 : 1);   // wave_skip: the persistent kernel may not have shaded anything
        for (igb200_ctx::FsIter& f : c->fs_inflight) f.shades_left -= shades;
        for (int k = 0; k < n_iter; ++k) c->fs_inflight.push_back(igb200_ctx::FsIter{st->iter + k, (long long)std::max(c->dev.max_depth, 1)});
        if (defer == 0) for (igb200_ctx::FsIter& f : c->fs_inflight) f.shades_left = 0;   // the launch ran every path to its end
        { const int r = fs_publish(c, false); if (r) return r; }
    }
    if (rays) { const int r = sync_control(c); if (r) return r; }   // the caller may free `rays` after the call
    return 0;
}

int igb200_render(igb200_ctx* c, const igb200_settings* st, const igb200_ray* rays, size_t n_rays) {
    if (!c || !st) return fail(-1, "igb200_render: null argument");
    if (!c->has_scene) return fail(-1, "igb200_render: no scene assigned");
    if (st->spi < 1) return fail(-1, "igb200_render: spi must be >= 1");
    if (rays) {   // igtrace: synchronous, never fused
        if (n_rays < 1 || (long long)n_rays * st->spi >= ((long long)1 << 31)) return fail(-1, "igb200_render: %zu rays at %d samples per iteration do not fit the 32-bit ray id", n_rays, st->spi);
        { const int r = flush_queued(c); if (r) return r; }
        return launch_iterations(c, st, 1, rays, n_rays);
    }
    if (st->width < 1 || st->height < 1) return fail(-1, "igb200_render: invalid size %dx%d", st->width, st->height);
    // the ray id (y * width + x) * spi + sample is a 32-bit integer in the stream, as in the reference (driver/mapping_cpu.art:351)
    if ((long long)st->width * st->height * st->spi >= ((long long)1 << 31)) return fail(-1, "igb200_render: %dx%d at %d samples per iteration overflows the 32-bit ray id", st->width, st->height, st->spi);
    // Fused iterations: a launch generates the camera rays of up to F consecutive iterations (same settings, iter + 1 each).
    if (!c->queued.empty()) {
        const igb200_settings& l = c->queued.back();
        const bool follows = l.spi == st->spi && l.width == st->width && l.height == st->height && l.frame == st->frame && l.seed == st->seed &&
                             l.device == st->device && st->iter == l.iter + 1;
        if (!follows) { const int r = flush_queued(c); if (r) return r; }
    }
    c->queued.push_back(*st);
    if ((int)c->queued.size() >= fuse_factor(c, st)) return flush_queued(c);
    return 0;
}

// Runs the stand-alone trace phase over n imported rays. any_hit: rays go through the shadow queue and the fused splat
// (unoccluded ray i adds 1 to fb[3 i]); else through the primary queue.
static int run_trace(igb200_ctx* c, const igb200_ray* d_rays, const uint32_t* d_flags, size_t n, int any_hit, float* d_fb, int repeat, double* ms_per_pass) {
    { const int r = sync_control(c); if (r) return r; }   // the hooks borrow the render queues
    { const int r = ensure_queues(c, n); if (r) return r; }
    if (n > c->capacity) return fail(-1, "igb200_trace_*: %zu rays exceed the queue capacity %zu", n, c->capacity);
    const int tb = c->trace_blocks ? c->trace_blocks : c->min_blocks;
    const bool flat = c->flat_on && tb < 4;
    const TraceKernel tk = trace_kernel(tb, c->vote, c->stage_where, flat);
    const size_t smem = flat ? c->smem_flat : c->smem_bytes;
    CU(cudaFuncSetAttribute((const void*)tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)tk, WF_BLOCK, smem));
    if (nb < 1) return fail(-2, "k_trace does not fit an SM with %zu bytes of shared memory", c->smem_bytes);
    const int grid = nb * c->n_sm;
    k_import_rays<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(d_rays, d_flags, any_hit ? RAY_SHADOW : RAY_CAMERA, (int)n, c->qa.view(), ShadowQueue{c->sq_org.p, c->sq_dir.p, c->sq_col.p}, any_hit);
    for (int r = 0; r < repeat + (ms_per_pass ? 3 : 0); ++r) {
        if (ms_per_pass && r == 3) CU(cudaEventRecord(c->ev0, c->stream));
        CU(cudaMemsetAsync(c->control.p, 0, CONTROL_SCRATCH, c->stream));
        if (!any_hit && r > 0) k_import_rays<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(d_rays, d_flags, RAY_CAMERA, (int)n, c->qa.view(), ShadowQueue{c->sq_org.p, c->sq_dir.p, c->sq_col.p}, 0);
        tk<<<grid, WF_BLOCK, smem, c->stream>>>(c->dev, c->qa.view(), any_hit ? 0 : (int)n, ShadowQueue{c->sq_org.p, c->sq_dir.p, c->sq_col.p}, any_hit ? (int)n : 0, d_fb,
                                                                                     &c->control.p->fetch_trace, &c->control.p->cam_launch, flat ? c->dev.n_flat_nodes : c->stage_nodes, c->stage_tris, c->stage_ent, c->refill,
                                                                                     (int)std::min<int64_t>(c->wide_rays_per_group * grid * (WF_BLOCK / 8), (int64_t)1 << 30));
    }
    if (ms_per_pass) {
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaEventSynchronize(c->ev1));
        float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *ms_per_pass = ms / repeat;
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

static int trace_list(igb200_ctx* c, const igb200_ray* rays, const uint32_t* flags, size_t n, int any, igb200_hit* out, int32_t* occ, int repeat, double* ms_per_pass) {
    if (!c || !rays) return fail(-1, "igb200_trace_*: null argument");
    if (!c->has_scene) return fail(-1, "igb200_trace_*: no scene assigned");
    CU(cudaSetDevice(c->device));
    if (n == 0) return 0;
    if (n > (size_t)1 << 30) return fail(-1, "igb200_trace_*: too many rays");
    DevBuf<igb200_ray> dr; DevBuf<uint32_t> df; DevBuf<igb200_hit> dh; DevBuf<float> dfb;
    CU(dr.alloc(n)); CU(cudaMemcpyAsync(dr.p, rays, n * sizeof(igb200_ray), cudaMemcpyHostToDevice, c->stream));
    if (flags) { CU(df.alloc(n)); CU(cudaMemcpyAsync(df.p, flags, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream)); }
    if (any) { CU(dfb.alloc(3 * n)); CU(cudaMemsetAsync(dfb.p, 0, 3 * n * sizeof(float), c->stream)); } else CU(dh.alloc(n));
    { const int r = run_trace(c, dr.p, df.p, n, any, dfb.p, repeat, ms_per_pass); if (r) return r; }
    if (any && occ) {
        std::vector<float> h(3 * n);
        CU(cudaMemcpy(h.data(), dfb.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) occ[i] = h[3 * i] == 0.0f;
    } else if (!any && out) {
        k_export_hits<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(c->qa.view(), (int)n, dh.p);
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        CU(cudaMemcpy(out, dh.p, n * sizeof(igb200_hit), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int igb200_trace_closest(igb200_ctx* c, const igb200_ray* rays, const uint32_t* flags, size_t n, igb200_hit* out) {
    if (!out) return fail(-1, "igb200_trace_closest: null argument");
    return trace_list(c, rays, flags, n, 0, out, nullptr, 1, nullptr);
}
int igb200_trace_any(igb200_ctx* c, const igb200_ray* rays, size_t n, int32_t* occluded) {
    if (!occluded) return fail(-1, "igb200_trace_any: null argument");
    return trace_list(c, rays, nullptr, n, 1, nullptr, occluded, 1, nullptr);
}
int igb200_bench_trace(igb200_ctx* c, const igb200_ray* rays, size_t n, int any_hit, int repeat, double* ms_per_pass) {
    if (!ms_per_pass || repeat < 1) return fail(-1, "igb200_bench_trace: bad argument");
    return trace_list(c, rays, nullptr, n, any_hit, nullptr, nullptr, repeat, ms_per_pass);
}

// ---- multi-GPU exchange: NCCL point-to-point over NVLink, loaded at run time -------------------------------------------------------
namespace {
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr; decltype(&ncclCommInitRank) CommInitRank = nullptr; decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr; decltype(&ncclRecv) Recv = nullptr; decltype(&ncclGroupStart) GroupStart = nullptr; decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr; decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string where;
};
NcclApi g_nccl;
int load_nccl() {
    if (g_nccl.handle) return 0;
    // the copy the process already has (PyTorch brings its own) wins: two NCCL builds in one process must not be mixed
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    std::string where = "libnccl.so.2 (already loaded in the process)";
    if (!h) { const char* env = getenv("IGB200_NCCL_LIB"); if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_GLOBAL); where = env; } }
    if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); where = "libnccl.so.2 (system)"; }
    if (!h) return fail(-3, "igb200_comm: cannot load libnccl.so.2 (%s); set IGB200_NCCL_LIB", dlerror());
    NcclApi a; a.handle = h; a.where = where;
#define IGB_SYM(N) a.N = reinterpret_cast<decltype(a.N)>(dlsym(h, "nccl" #N)); if (!a.N) return fail(-3, "igb200_comm: nccl" #N " not found in %s", where.c_str());
    IGB_SYM(GetUniqueId) IGB_SYM(CommInitRank) IGB_SYM(CommDestroy) IGB_SYM(Send) IGB_SYM(Recv) IGB_SYM(GroupStart) IGB_SYM(GroupEnd) IGB_SYM(GetErrorString) IGB_SYM(GetVersion)
#undef IGB_SYM
    g_nccl = a;
    return 0;
}
}  // namespace
#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(-2, "%s failed: %s (%s:%d)", #x, g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)

int igb200_comm_unique_id(uint8_t id[128]) {
    if (!id) return fail(-1, "igb200_comm_unique_id: null argument");
    { const int r = load_nccl(); if (r) return r; }
    ncclUniqueId u;
    static_assert(sizeof(u) == 128, "ncclUniqueId");
    NC(g_nccl.GetUniqueId(&u));
    std::memcpy(id, &u, 128);
    return 0;
}

int igb200_comm_init(igb200_ctx* c, int rank, int world, int tile_size, const uint8_t id[128]) {
    if (!c || !id) return fail(-1, "igb200_comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world || tile_size < 1) return fail(-1, "igb200_comm_init: invalid rank %d / world %d / tile %d", rank, world, tile_size);
    if (c->comm) return fail(-1, "igb200_comm_init: this context already has a communicator");
    { const int r = load_nccl(); if (r) return r; }
    { const int r = igb200_set_partition(c, rank, world, tile_size); if (r) return r; }
    CU(cudaSetDevice(c->device));
    ncclUniqueId u; std::memcpy(&u, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, world, u, rank));
    return 0;
}

int igb200_comm_destroy(igb200_ctx* c) {
    if (!c) return fail(-1, "null context");
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->comm);
        c->comm = nullptr;
    }
    if (c->comm_host) { cudaFreeHost(c->comm_host); c->comm_host = nullptr; c->comm_host_n = 0; }
    return 0;
}

// Tables and buffers of the exchange for the current frame size
static int comm_prepare(igb200_ctx* c) {
    const int W = c->width, H = c->height, tile = c->tile, world = c->world, rank = c->rank;
    const int tiles_x = (W + tile - 1) / tile, tiles_y = (H + tile - 1) / tile;
    const long long per_tile = (long long)tile * tile * 3;
    const long long key[4] = {W, H, tile, world};
    if (std::memcmp(key, c->comm_key, sizeof(key)) != 0) {
        std::vector<long long> n_tiles(world, 0), offs(world, 0);
        std::vector<int> local_index((size_t)tiles_x * tiles_y);
        for (int t = 0; t < tiles_x * tiles_y; ++t) { const int r = (t % tiles_x + t / tiles_x) % world; local_index[t] = (int)n_tiles[r]++; }
        long long total = 0;
        for (int r = 0; r < world; ++r) { offs[r] = total; total += n_tiles[r] * per_tile; }
        c->comm_counts.assign(world, 0);
        for (int r = 0; r < world; ++r) c->comm_counts[r] = n_tiles[r] * per_tile;
        CU(cudaStreamSynchronize(c->stream));
        CU(c->comm_send.alloc((size_t)c->comm_counts[rank]));
        if (rank == 0) {
            CU(c->comm_recv.alloc((size_t)total));
            CU(c->comm_frame.alloc((size_t)W * H * 3));
            CU(c->comm_local_index.upload(local_index));
            CU(c->comm_rank_offset.upload(offs));
        }
        std::memcpy(c->comm_key, key, sizeof(key));
    }
    return ensure_tile_table(c, W, H);
}
// The exchange itself, asynchronous on the context's stream: the tiles this rank owns out of `src` -> rank 0, which assembles `frame`
static int comm_gather_async(igb200_ctx* c, const float* src, float* frame) {
    NvtxRange nvtx_("igb200_comm: gather tiles (NCCL)");
    { const int r = comm_prepare(c); if (r) return r; }
    const int W = c->width, H = c->height, tile = c->tile, world = c->world, rank = c->rank;
    const int tiles_x = (W + tile - 1) / tile;
    const int grid = c->n_sm * 8;
    std::vector<long long> offs(world, 0);
    { long long t = 0; for (int r = 0; r < world; ++r) { offs[r] = t; t += c->comm_counts[r]; } }
    if (world == 1) { CU(cudaMemcpyAsync(frame, src, (size_t)W * H * 3 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream)); return 0; }
    // rank 0 packs straight into its region of the receive buffer; the others into their send buffer
    float* pack_to = rank == 0 ? c->comm_recv.p + offs[0] : c->comm_send.p;
    k_pack_tiles<<<grid, 256, 0, c->stream>>>(src, pack_to, c->tile_table.p, (int)c->n_local_tiles, tile, tiles_x, W, H);
    NC(g_nccl.GroupStart());
    if (rank == 0) { for (int r = 1; r < world; ++r) NC(g_nccl.Recv(c->comm_recv.p + offs[r], (size_t)c->comm_counts[r], ncclFloat32, r, c->comm, c->stream)); }
    else NC(g_nccl.Send(c->comm_send.p, (size_t)c->comm_counts[rank], ncclFloat32, 0, c->comm, c->stream));
    NC(g_nccl.GroupEnd());
    if (rank == 0) k_unpack_tiles<<<grid, 256, 0, c->stream>>>(frame, c->comm_recv.p, c->comm_local_index.p, c->comm_rank_offset.p, world, tile, tiles_x, W, H);
    CU(cudaGetLastError());
    c->pending = true;
    return 0;
}

int igb200_comm_gather_framebuffer(igb200_ctx* c, const char* aov, float** device_frame, float** host_frame) {
    if (!c) return fail(-1, "null context");
    if (!c->comm) return fail(-1, "igb200_comm_gather_framebuffer: no communicator (igb200_comm_init)");
    const int which = aov_index(c, aov);
    if (which < 0) return fail(-4, "igb200_comm_gather_framebuffer: AOV '%s' does not exist", aov);
    { const int r = flush_queued(c); if (r) return r; }
    if (!c->fb.p) return fail(-1, "igb200_comm_gather_framebuffer: no framebuffer");
    CU(cudaSetDevice(c->device));
    { const int r = drain(c); if (r) return r; }   // asynchronous: the exchange below is ordered behind it on the stream
    const float* src = which == 0 ? c->fb.p : c->aov[which - 1].p;
    if (!src) return fail(-1, "igb200_comm_gather_framebuffer: AOV '%s' is not allocated", aov);
    { const int r = comm_prepare(c); if (r) return r; }
    { const int r = comm_gather_async(c, src, c->comm_frame.p); if (r) return r; }
    if (device_frame) *device_frame = nullptr;
    if (host_frame) *host_frame = nullptr;
    if (c->rank != 0) return 0;
    if (device_frame) *device_frame = c->comm_frame.p;
    if (host_frame) {
        const size_t n = (size_t)c->width * c->height * 3;
        if (c->comm_host_n != n) {
            if (c->comm_host) { cudaFreeHost(c->comm_host); c->comm_host = nullptr; }
            CU(cudaMallocHost(&c->comm_host, n * sizeof(float)));
            c->comm_host_n = n;
        }
        CU(cudaMemcpyAsync(c->comm_host, c->comm_frame.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        *host_frame = c->comm_host;
    }
    return 0;
}

// ---- frame streaming -----------------------------------------------------------------------------------------------------------
static int fs_host_buffer(igb200_ctx* c, int* out) {
    if (!c->fs_host_free.empty()) { *out = c->fs_host_free.back(); c->fs_host_free.pop_back(); return 0; }
    float* p = nullptr;
    CU(cudaMallocHost(&p, c->fb.n * sizeof(float)));
    c->fs_host.push_back(p);
    *out = (int)c->fs_host.size() - 1;
    return 0;
}
// Publishes finished iterations in order (all of them when `all`: the caller has just finished every path). Asynchronous.
static int fs_publish(igb200_ctx* c, bool all) {
    NvtxRange nvtx_("igb200: publish frames");
    const size_t n = c->fb.n;
    const bool root = c->rank == 0;
    while (!c->fs_inflight.empty() && (all || c->fs_inflight.front().shades_left <= 0)) {
        const igb200_ctx::FsIter f = c->fs_inflight.front();
        c->fs_inflight.pop_front();
        const int p = (int)(c->fs_published & 1);
        float* slot = c->fs_ring.p + (size_t)(f.iter & (c->fs_slots - 1)) * n;
        float* snap = nullptr;
        if (c->fs_shm) {
            // ---- frames in shared host memory: no exchange between the GPUs at all. Every rank snapshots its accumulated frame and writes its own
            // tiles into host frame h of the segment on its copy stream, then raises flags[h][rank] = sequence number of the frame.
            const unsigned int seq = (unsigned int)(c->fs_published + 1);
            const int h = (int)(c->fs_published % c->fs_shm_frames);
            volatile unsigned long long* consumed = reinterpret_cast<volatile unsigned long long*>(c->fs_shm);
            if ((unsigned long long)c->fs_published - *consumed >= (unsigned long long)c->fs_shm_frames) {
                // host frame h still belongs to a frame rank 0's caller has not taken (and released) yet
                if (root) return fail(-5, "frame stream: %d published frames are waiting to be taken (igb200_frame_stream_next); none can be overwritten", c->fs_shm_frames);
                const auto t0 = std::chrono::steady_clock::now();
                while ((unsigned long long)c->fs_published - *consumed >= (unsigned long long)c->fs_shm_frames)
                    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0) return fail(-5, "frame stream: rank 0 has not taken a frame for 60 s");
            }
            if (c->fs_snap_used[p]) CU(cudaStreamWaitEvent(c->stream, c->fs_ev_snapfree[p], 0));
            snap = c->fs_snap[p].p;
            k_publish<<<c->n_sm * 4, 256, 0, c->stream>>>(c->fb.p, slot, snap, (long long)n);
            CU(cudaEventRecord(c->fs_ev_pub[p], c->stream));
            CU(cudaStreamWaitEvent(c->copy_stream, c->fs_ev_pub[p], 0));
            float* host_frame = reinterpret_cast<float*>(c->fs_shm_dev + c->fs_shm_header + (size_t)h * c->fs_shm_frame_bytes);
            const int tiles_x = (c->width + c->tile - 1) / c->tile;
            if ((c->tile * 3) % 4 == 0 && (c->width * 3) % 4 == 0) k_tiles_to_host<true><<<c->n_sm * 4, 256, 0, c->copy_stream>>>(snap, host_frame, c->tile_table.p, (int)c->n_local_tiles, c->tile, tiles_x, c->width, c->height);
            else k_tiles_to_host<false><<<c->n_sm * 4, 256, 0, c->copy_stream>>>(snap, host_frame, c->tile_table.p, (int)c->n_local_tiles, c->tile, tiles_x, c->width, c->height);
            k_raise_flag<<<1, 1, 0, c->copy_stream>>>(reinterpret_cast<volatile unsigned int*>(c->fs_shm_dev + 64) + (size_t)h * 64 + c->rank, seq);
            CU(cudaEventRecord(c->fs_ev_snapfree[p], c->copy_stream));
            CU(cudaGetLastError());
            c->fs_snap_used[p] = true;
            c->fs_published++;
            c->pending = true;
            if (root) c->fs_ready.push_back(igb200_ctx::FsFrame{f.iter, h, nullptr, seq});
            continue;
        }
        if (root) {
            // the snapshot the copy engine reads must not be overwritten before its previous copy is through
            if (c->fs_snap_used[p]) CU(cudaStreamWaitEvent(c->stream, c->fs_ev_snapfree[p], 0));
            snap = c->fs_snap[p].p;
        }
        const bool gather = c->comm != nullptr && c->world > 1;
        k_publish<<<c->n_sm * 4, 256, 0, c->stream>>>(c->fb.p, slot, gather ? nullptr : snap, (long long)n);
        if (gather) { const int r = comm_gather_async(c, c->fb.p, snap); if (r) return r; }   // every rank's tiles of the frame -> rank 0's snapshot
        CU(cudaGetLastError());
        c->fs_published++;
        c->pending = true;
        if (!root) continue;
        CU(cudaEventRecord(c->fs_ev_pub[p], c->stream));
        int h = -1;
        { const int r = fs_host_buffer(c, &h); if (r) return r; }
        CU(cudaStreamWaitEvent(c->copy_stream, c->fs_ev_pub[p], 0));
        CU(cudaMemcpyAsync(c->fs_host[h], snap, n * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaEventRecord(c->fs_ev_snapfree[p], c->copy_stream));
        c->fs_snap_used[p] = true;
        cudaEvent_t e;
        if (!c->fs_events.empty()) { e = c->fs_events.back(); c->fs_events.pop_back(); } else CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventRecord(e, c->copy_stream));
        c->fs_ready.push_back(igb200_ctx::FsFrame{f.iter, h, e, 0u});
    }
    return 0;
}

int igb200_frame_stream_begin(igb200_ctx* c, int slots) {
    if (!c) return fail(-1, "null context");
    if (c->fs_on) return fail(-1, "igb200_frame_stream_begin: already streaming");
    if (c->deterministic) return fail(-1, "igb200_frame_stream_begin: frame streaming and deterministic accumulation exclude each other");
    if (!c->fb.p) return fail(-1, "igb200_frame_stream_begin: no framebuffer (call igb200_resize first)");
    if ((long long)c->width * c->height >= (1ll << 24)) return fail(-4, "igb200_frame_stream_begin: frames of 2^24 pixels and more are not supported (the slot rides in the shadow ray's pixel word)");
    if (slots <= 0) slots = 16;
    int R = 2; while (R < slots) R <<= 1;
    if (R > 128) return fail(-1, "igb200_frame_stream_begin: at most 128 slots");
    { const int r = sync_control(c); if (r) return r; }   // nothing in flight when the framebuffer layout changes
    CU(cudaSetDevice(c->device));
    const size_t n = c->fb.n;
    CU(c->fs_ring.alloc((size_t)R * n));
    CU(cudaMemset(c->fs_ring.p, 0, (size_t)R * n * sizeof(float)));
    const bool shared = c->fs_share_key != 0 && c->world > 1;
    if (shared) {
        { const int r = ensure_tile_table(c, c->width, c->height); if (r) return r; }
        if (c->world > 64) return fail(-1, "igb200_frame_stream_share: at most 64 ranks");
        c->fs_shm_frames = 2 * R;
        c->fs_shm_frame_bytes = (n * sizeof(float) + 4095) & ~(size_t)4095;
        c->fs_shm_header = (64 + (size_t)c->fs_shm_frames * 64 * sizeof(unsigned int) + 4095) & ~(size_t)4095;
        c->fs_shm_bytes = c->fs_shm_header + (size_t)c->fs_shm_frames * c->fs_shm_frame_bytes;
        // rank 0 creates the segment, the others attach once it exists with the right size
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            c->fs_shm_id = c->rank == 0 ? shmget((key_t)c->fs_share_key, c->fs_shm_bytes, IPC_CREAT | IPC_EXCL | 0600) : shmget((key_t)c->fs_share_key, c->fs_shm_bytes, 0600);
            if (c->fs_shm_id >= 0) break;
            if (c->rank == 0 && errno == EEXIST && !c->fs_shm_retry) {   // left behind by a run that died: remove it and try once more
                const int old_id = shmget((key_t)c->fs_share_key, 0, 0600);
                if (old_id >= 0) shmctl(old_id, IPC_RMID, nullptr);
                c->fs_shm_retry = true;
                continue;
            }
            if (c->rank == 0) return fail(-5, "igb200_frame_stream_begin: cannot create the shared segment (key %d, %zu bytes): %s", c->fs_share_key, c->fs_shm_bytes, strerror(errno));
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) return fail(-5, "igb200_frame_stream_begin: the shared segment (key %d) did not appear: %s", c->fs_share_key, strerror(errno));
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
        void* m = shmat(c->fs_shm_id, nullptr, 0);
        if (m == (void*)-1) { c->fs_shm_id = -1; return fail(-5, "igb200_frame_stream_begin: shmat: %s", strerror(errno)); }
        // (a new System V segment is zero-filled by the kernel: consumed = 0, no flag raised)
        void* dp = nullptr;
        cudaError_t e = cudaHostRegister(m, c->fs_shm_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
        if (e == cudaSuccess) { e = cudaHostGetDevicePointer(&dp, m, 0); if (e != cudaSuccess) cudaHostUnregister(m); }
        if (e != cudaSuccess) {   // leave nothing half set up: the caller may fall back to the gathered stream
            cudaGetLastError();
            shmdt(m);
            if (c->rank == 0) shmctl(c->fs_shm_id, IPC_RMID, nullptr);
            c->fs_shm_id = -1;
            return fail(-2, "igb200_frame_stream_begin: cannot pin the shared segment (%zu bytes): %s", c->fs_shm_bytes, cudaGetErrorString(e));
        }
        c->fs_shm = static_cast<unsigned char*>(m);
        c->fs_shm_dev = static_cast<unsigned char*>(dp);
        c->fs_taken_shared = false; c->fs_shm_retry = false;
    }
    if (c->rank == 0 || shared) for (int k = 0; k < 2; ++k) {
        CU(c->fs_snap[k].alloc(n));
        if (!c->fs_ev_pub[k]) { CU(cudaEventCreateWithFlags(&c->fs_ev_pub[k], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->fs_ev_snapfree[k], cudaEventDisableTiming)); }
        c->fs_snap_used[k] = false;
    }
    c->fs_slots = R; c->fs_on = true; c->fs_published = 0; c->fs_last_taken_host = -1;
    // pinned frames for everything that can be outstanding at once (a drain publishes every iteration in flight in one go); pinning
    // memory costs milliseconds per frame, so it happens here and not inside a render
    if (c->rank == 0 && !shared) {
        std::vector<int> hs((size_t)R + 2);
        for (int& h : hs) { const int r = fs_host_buffer(c, &h); if (r) return r; }
        for (int h : hs) c->fs_host_free.push_back(h);
    }
    return 0;
}

int igb200_frame_stream_share(igb200_ctx* c, int key) {
    if (!c) return fail(-1, "null context");
    if (c->fs_on) return fail(-1, "igb200_frame_stream_share: call before igb200_frame_stream_begin");
    if (key < 0) return fail(-1, "igb200_frame_stream_share: the key must be positive (0 = off)");
    c->fs_share_key = key;
    return 0;
}

// Releases the ring; frames not taken yet are dropped. The accumulated framebuffer keeps everything that was rendered.
int igb200_frame_stream_end(igb200_ctx* c) {
    if (!c) return fail(-1, "null context");
    if (!c->fs_on) return 0;
    { const int r = sync_control(c); if (r) return r; }
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    for (const igb200_ctx::FsFrame& f : c->fs_ready) if (f.copied) c->fs_events.push_back(f.copied);
    c->fs_ready.clear(); c->fs_inflight.clear();
    for (float* p : c->fs_host) cudaFreeHost(p);
    c->fs_host.clear(); c->fs_host_free.clear();
    c->fs_ring.release(); c->fs_snap[0].release(); c->fs_snap[1].release();
    if (c->fs_shm) {
        cudaHostUnregister(c->fs_shm);
        shmdt(c->fs_shm);
        if (c->rank == 0) shmctl(c->fs_shm_id, IPC_RMID, nullptr);   // gone once the last rank has detached
        c->fs_shm = nullptr; c->fs_shm_dev = nullptr; c->fs_shm_id = -1; c->fs_shm_frames = 0;
    }
    c->fs_on = false; c->fs_slots = 0;
    return 0;
}

int igb200_frame_stream_next(igb200_ctx* c, int wait, int* iteration, float** host_rgb) {
    if (!c || !iteration || !host_rgb) return fail(-1, "igb200_frame_stream_next: null argument");
    if (!c->fs_on) return fail(-1, "igb200_frame_stream_next: not streaming (igb200_frame_stream_begin)");
    *host_rgb = nullptr; *iteration = -1;
    CU(cudaSetDevice(c->device));
    if (c->fs_last_taken_host >= 0) { c->fs_host_free.push_back(c->fs_last_taken_host); c->fs_last_taken_host = -1; }   // the previous frame's buffer is the caller's no longer
    if (c->fs_taken_shared) { volatile unsigned long long* consumed = reinterpret_cast<volatile unsigned long long*>(c->fs_shm); *consumed = *consumed + 1; c->fs_taken_shared = false; }
    if (wait >= 2) { const int r = drain(c); if (r) return r; }   // finish everything that was rendered: every frame becomes ready
    else if (wait == 1 || c->fs_inflight.empty()) { const int r = flush_queued(c); if (r) return r; }   // the same decision on every rank
    // (a poll -- wait 0 -- leaves queued iterations alone while the device still has frames in the works: launching them one by one would
    // undo the fusing of small iterations, which is what a rank's share of a frame at N = 8 lives on: r6q, e2e 29 Grays/s with every step's
    // poll flushing the queue)
    if (c->rank != 0 || c->fs_ready.empty()) return 0;
    const igb200_ctx::FsFrame f = c->fs_ready.front();
    if (c->fs_shm) {
        // complete when every rank has raised its flag for this frame (each rank's tiles are in the segment by then)
        volatile unsigned int* flags = reinterpret_cast<volatile unsigned int*>(c->fs_shm + 64) + (size_t)f.host * 64;
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            bool all = true;
            for (int r = 0; r < c->world; ++r) all = all && flags[r] == f.seq;
            if (all) break;
            if (wait == 0) return 0;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0) return fail(-5, "igb200_frame_stream_next: frame of iteration %d did not arrive from every rank within 120 s", f.iter);
        }
        __sync_synchronize();
        c->fs_ready.pop_front();
        c->fs_taken_shared = true;
        *iteration = f.iter; *host_rgb = reinterpret_cast<float*>(c->fs_shm + c->fs_shm_header + (size_t)f.host * c->fs_shm_frame_bytes);
        return 1;
    }
    if (wait == 0) {
        const cudaError_t q = cudaEventQuery(f.copied);
        if (q == cudaErrorNotReady) return 0;
        CU(q);
    } else CU(cudaEventSynchronize(f.copied));
    c->fs_ready.pop_front();
    c->fs_events.push_back(f.copied);
    c->fs_last_taken_host = f.host;
    *iteration = f.iter; *host_rgb = c->fs_host[f.host];
    return 1;
}

// Test hook: builds a BVH8 over n boxes (6 floats each) with the host builder (builder 0, needs no GPU and no context) or the GPU builder
// (builder 1), optionally through the cache in `dir` (store, then load and compare), validates it -- every primitive in exactly one leaf,
// every child box contains what hangs below it -- and reports {nodes, depth, leaves, SAH cost x 1000 relative to the root's area}.
int igb200_test_bvh_build(igb200_ctx* c, const float* boxes6, size_t n, int builder, const char* dir, int64_t out[4]) {
    if (!boxes6 || !out || n < 1) return fail(-1, "igb200_test_bvh_build: bad argument");
    std::vector<Box3> boxes(n);
    std::memcpy(boxes.data(), boxes6, n * sizeof(Box3));
    Bvh8 bvh;
    if (builder == BVH_BUILDER_GPU_LBVH) {
        if (!c) return fail(-1, "igb200_test_bvh_build: the GPU builder needs a context");
        CU(cudaSetDevice(c->device));
        std::string err;
        if (!build_bvh8_gpu(boxes, bvh, c->stream, err)) return fail(-2, "%s", err.c_str());
    } else bvh = build_bvh8(boxes, 4);
    if (dir && *dir) {
        const uint64_t h = bvh_cache_hash(boxes, 4, builder);
        if (!bvh_cache_store(dir, h, builder, bvh)) return fail(-2, "igb200_test_bvh_build: cannot store into '%s'", dir);
        Bvh8 back;
        if (!bvh_cache_load(dir, h, n, builder, back)) return fail(-2, "igb200_test_bvh_build: cannot load what was stored");
        if (back.nodes.size() != bvh.nodes.size() || back.order != bvh.order || back.max_depth != bvh.max_depth ||
            std::memcmp(back.nodes.data(), bvh.nodes.data(), bvh.nodes.size() * sizeof(Node8)) != 0) return fail(-2, "igb200_test_bvh_build: the cached tree differs from the stored one");
        if (bvh_cache_load(dir, h ^ 1, n, builder, back)) return fail(-2, "igb200_test_bvh_build: a tree was loaded under a hash it was not stored under");
    }
    // validation: depth-first from the root with the box each subtree must stay inside
    std::vector<int> seen(n, 0);
    struct Item { int node; Box3 box; int depth; };
    std::vector<Item> stack;
    Box3 root = Box3::empty();
    for (const Box3& b : boxes) root.extend(b);
    stack.push_back(Item{0, root, 1});
    auto inside = [](const Box3& a, const Box3& b) { for (int k = 0; k < 3; ++k) if (a.lo[k] < b.lo[k] || a.hi[k] > b.hi[k]) return false; return true; };
    double cost = 0; int64_t leaves = 0; int depth = 0;
    if (bvh.nodes.empty()) return fail(-2, "igb200_test_bvh_build: empty tree");
    while (!stack.empty()) {
        const Item it = stack.back(); stack.pop_back();
        depth = std::max(depth, it.depth);
        if (it.node < 0 || (size_t)it.node >= bvh.nodes.size()) return fail(-2, "igb200_test_bvh_build: child index %d out of range", it.node);
        const Node8& nd = bvh.nodes[it.node];
        bool ended = false;
        for (int k = 0; k < 8; ++k) {
            const int ch = nd.child[k];
            if (ch == 0) { ended = true; continue; }
            if (ended) return fail(-2, "igb200_test_bvh_build: node %d: children are not packed to the front", it.node);
            Box3 cb; for (int a = 0; a < 3; ++a) { cb.lo[a] = nd.bounds[2 * a][k]; cb.hi[a] = nd.bounds[2 * a + 1][k]; }
            if (!inside(cb, it.box)) return fail(-2, "igb200_test_bvh_build: node %d child %d sticks out of its parent's box", it.node, k);
            cost += cb.half_area();
            if (ch > 0) stack.push_back(Item{ch - 1, cb, it.depth + 1});
            else {
                const int r = -ch - 1, first = r >> 2, cnt = (r & 3) + 1;
                ++leaves;
                for (int j = 0; j < cnt; ++j) {
                    if ((size_t)(first + j) >= n) return fail(-2, "igb200_test_bvh_build: leaf slot %d out of range", first + j);
                    const int p = bvh.order[first + j];
                    if (p < 0 || (size_t)p >= n || seen[p]++) return fail(-2, "igb200_test_bvh_build: primitive %d is referenced twice or out of range", p);
                    if (!inside(boxes[p], cb)) return fail(-2, "igb200_test_bvh_build: primitive %d sticks out of its leaf's box", p);
                }
            }
        }
    }
    for (size_t p = 0; p < n; ++p) if (!seen[p]) return fail(-2, "igb200_test_bvh_build: primitive %zu is in no leaf", p);
    if (depth != bvh.max_depth && n > 4) return fail(-2, "igb200_test_bvh_build: max_depth says %d, the tree has %d levels", bvh.max_depth, depth);
    out[0] = (int64_t)bvh.nodes.size(); out[1] = depth; out[2] = leaves; out[3] = (int64_t)(1000.0 * cost / std::max((double)root.half_area(), 1e-30));
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
