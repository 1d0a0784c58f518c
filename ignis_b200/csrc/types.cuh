// Device-side data layout of the wavefront path tracer (DESIGN.md "Data layout in HBM").
//
// Rays live in HBM as structure-of-float4-arrays queues, so that a warp's load of one field of 32 consecutive
// records is four fully used 128-byte lines; survivors are appended to the next queue with one atomic per warp
// (ballot + popc prefix), which replaces the reference's sort / compact kernels and their host round trips
// (driver/mapping_gpu.art:409-502,686-711). BVH8 nodes are two 128-byte lines. Tensor cores are not used: there is
// no dense contraction anywhere on this path.
#pragma once

#include "device_math.cuh"

namespace igb {

struct DevScene {
    const float4* nodes;      // BVH8 nodes, 16 float4 each (top level first, then one tree per shape)
    const float4* tris;       // 3 float4 per primitive slot: (v0,n.x) (e1,n.y) (e2,n.z)
    const int*    tri_prim;   // primitive id of each slot
    const float4* ent_leaf;   // 8 float4 per entity: see upload in api.cu
    const float4* ent_shade;  // 6 float4 per entity: global rows 0-2, normal rows 0-2 (.w = shape_id, mat_id, 0)
    const float4* blob;       // `shapes` dyn-table data
    const int4*   shape_info; // 2 int4 per shape: (type, v_start, n_start, i_start) (tex_start_f2, n_face, 0, 0)
    const float4* materials;  // 8 float4 per material (igb200_material, 128 bytes)
    const float*  textures;   // 24 words per texture (igb200_texture)
    const int4*   images;     // per image: format, width, height, first 32-bit word of its pixels inside image_data (16-byte aligned)
    const uint32_t* image_data;
    const float*  aux_data;   // 2-D cdfs of the textured environment lights
    const float*  inf_lights; // 32 words per light (igb200_light)
    const float*  fin_lights;
    int   n_ent, n_mat, n_inf, n_fin;
    int   n_nodes, n_tris;    // array sizes (for staging into shared memory)
    // Small scenes: ONE merged tree (api.cu build_flat_tree) -- the top-level tree with every instance's tree hanging off it, the
    // instances' node boxes refitted in WORLD space. Leaves carry the entity (traverse.cuh flat_leaf_step); triangles stay per shape.
    const float4* flat_nodes; int n_flat_nodes;
    float scene_radius;
    int   max_depth, min_depth;
    float clamp_value;
    int   nee;
    int   full;               // the launch needs the shade kernels with every feature (shade.cuh shade_record<true>): set per launch by api.cu make_params
    int   selector;           // light selector: 0 uniform, 1 flux cdf ("simple"), 2 light hierarchy (igb200_technique.light_selector)
    const float* selector_data;   // light_cdf.bin / light_hierarchy.bin words
    float eye[3], view[9];    // view = columns right, up, dir
    float scale_x, scale_y, cam_tmin, cam_tmax;
};

struct RenderParams {
    int   spi, iter, frame, seed, width, height;
    float inv_spi;
    int   tile_w, tile_h, tiles_x, rank, world;
    const int* tile_table;   // multi-GPU: row-major index of the rank's k-th tile (api.cu launch_iterations); null = all tiles
    float* aov_normals; float* aov_albedo;   // standard AOVs (technique/internal/infobuffer.art): written by camera-ray hits of iteration 0; null = off
    long long per_iter;   // camera-ray domain of ONE iteration; a launch may generate several consecutive iterations (api.cu: fused iterations)
    // Frame streaming (api.cu igb200_frame_stream_*): every iteration splats into its own slot `iter & ring_mask` of a ring of framebuffers
    // (ring_stride floats apart), so that a frame can be handed out while later iterations are already in flight. 0 / 0: one framebuffer.
    int ring_mask; long long ring_stride;
    // Deterministic accumulation (api.cu option "deterministic"): contributions go to a slot per SAMPLE (index = ray id) instead of the
    // pixel, so no two threads ever add to the same word; k_resolve then folds the slots into the frame in sample order.
    int det;
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

struct HitR { float t, u, v; int prim, ent; };
struct C3 { float r, g, b; };
__device__ __forceinline__ C3 c3(float r, float g, float b) { C3 c; c.r = r; c.g = g; c.b = b; return c; }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

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

// Adds the warp's total of `n` to a 64-bit statistics counter with one atomic (all 32 lanes must arrive).
__device__ __forceinline__ void warp_count(unsigned long long* counter, int n) {
    const int total = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(counter, (unsigned long long)total);
}

__device__ __forceinline__ void splat(float* fb, int pixel, C3 c, float inv_spi) {   // driver/accumulator.art:4-21
    atomicAdd(fb + (size_t)pixel * 3 + 0, c.r * inv_spi);
    atomicAdd(fb + (size_t)pixel * 3 + 1, c.g * inv_spi);
    atomicAdd(fb + (size_t)pixel * 3 + 2, c.b * inv_spi);
}

}  // namespace igb
