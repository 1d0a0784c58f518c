// BVH8 construction ON THE GPU + the on-disk BVH cache.
//
// What it replaces in the reference: the CPU build of every mesh's BVH through the external madmann91/bvh builder and its N-ary
// collapse (src/runtime/shape/TriMeshProvider.cpp:255-298 build_bvh, src/runtime/bvh/TriBVHAdapter.h:196-223,
// src/runtime/bvh/NArityBvh.h:94-143), and the cache that exists because that build dominates the load time of large meshes
// (TriMeshProvider.cpp:326-351: meshes of more than 500 000 faces are serialised to `<cache>/_bvh_<name>.bin`, keyed by a hash of the
// mesh).
//
// Build (one stream, no host round trip except one counter per tree level):
//   1. centroid bounds           k_centroid_bounds   warp min / max -> 6 atomics on order-preserving integer keys
//   2. 30-bit Morton codes       k_morton
//   3. sort (code, primitive)    cub::DeviceRadixSort (the one library call: a sort primitive of the CUDA toolkit, not on the render path)
//   4. binary radix tree         k_radix_tree        Karras 2012: every inner node finds its range and split from the sorted codes, in parallel
//   5. boxes bottom-up           k_fit               one thread per leaf climbs; the second arrival at a node merges the children
//   6. collapse to BVH8          k_collapse          level by level: a thread owns one wide node, opens the child with the largest surface
//                                                    area until eight are held; subtrees of <= 4 primitives become leaves (contiguous in
//                                                    Morton order, so a leaf is a (first slot, count) code as in bvh8.h)
// The output is the same Bvh8 (nodes, order, max_depth) the host builder makes, so everything downstream -- rebasing, merged tree, node
// order, upload -- is unchanged, and because the closest hit is a pure function of the ray (traverse.cuh header) renders are bit-identical
// whichever builder made the tree (tests/test_gpu_parity.py::test_gpu_built_bvh_is_invisible). An LBVH is a worse tree than the
// binned-SAH one (measured: DESIGN.md); it is for meshes whose host build would dominate the load.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <unistd.h>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bvh8.h"
#include "bvh_build.h"

namespace igb {
namespace {

__device__ __forceinline__ unsigned ordered(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float unordered(unsigned k) { const unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; float f; memcpy(&f, &u, 4); return f; }

// boxes: 6 floats per primitive (lo xyz, hi xyz). keys[0..2] = min, keys[3..5] = max of the centroids (ordered integer keys)
__global__ void k_centroid_bounds(const float* __restrict__ boxes, int n, unsigned* __restrict__ keys) {
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* b = boxes + 6 * (size_t)i;
        for (int k = 0; k < 3; ++k) { const unsigned c = ordered(0.5f * (b[k] + b[3 + k])); mn[k] = min(mn[k], c); mx[k] = max(mx[k], c); }
    }
    for (int k = 0; k < 3; ++k) {
        const unsigned a = __reduce_min_sync(0xffffffffu, mn[k]), b = __reduce_max_sync(0xffffffffu, mx[k]);
        if ((threadIdx.x & 31) == 0) { atomicMin(keys + k, a); atomicMax(keys + 3 + k, b); }
    }
}

__device__ __forceinline__ unsigned spread3(unsigned v) {   // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void k_morton(const float* __restrict__ boxes, int n, const unsigned* __restrict__ keys, unsigned* __restrict__ codes, int* __restrict__ prims) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + 6 * (size_t)i;
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        const float lo = unordered(keys[k]), hi = unordered(keys[3 + k]);
        const float ext = hi - lo;
        const float t = ext > 0 ? (0.5f * (b[k] + b[3 + k]) - lo) / ext : 0.0f;
        q[k] = (unsigned)fminf(fmaxf(t * 1024.0f, 0.0f), 1023.0f);
    }
    codes[i] = (spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2]);
    prims[i] = i;
}

// Length of the common prefix of the sorted keys i and j, ties broken by the index (Karras 2012, section 4); -1 outside the array
__device__ __forceinline__ int delta(const unsigned* __restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned a = codes[i], b = codes[j];
    return a == b ? 32 + __clz((unsigned)i ^ (unsigned)j) : __clz(a ^ b);
}

// Child references of the binary tree: >= 0 inner node, < 0 leaf ~ref (leaf j = the j-th primitive in Morton order)
__global__ void k_radix_tree(const unsigned* __restrict__ codes, int n, int2* __restrict__ children, int2* __restrict__ range, int* __restrict__ inner_parent, int* __restrict__ leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(codes, n, i, i - d);
    int lmax = 2;
    while (delta(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = first == gamma ? ~gamma : gamma, right = last == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    range[i] = make_int2(first, last);
    if (left >= 0) inner_parent[left] = i; else leaf_parent[~left] = i;
    if (right >= 0) inner_parent[right] = i; else leaf_parent[~right] = i;
    if (i == 0) inner_parent[0] = -1;
}

struct Fbox { float lo[3], hi[3]; };
__device__ __forceinline__ Fbox load_box(const float* p) { Fbox b; for (int k = 0; k < 3; ++k) { b.lo[k] = p[k]; b.hi[k] = p[3 + k]; } return b; }
__device__ __forceinline__ Fbox child_box(int ref, const float* __restrict__ boxes, const int* __restrict__ prims, const float* __restrict__ inner) {
    return ref >= 0 ? load_box(inner + 6 * (size_t)ref) : load_box(boxes + 6 * (size_t)prims[~ref]);
}

__global__ void k_fit(int n, const float* __restrict__ boxes, const int* __restrict__ prims, const int2* __restrict__ children, const int* __restrict__ inner_parent,
                      const int* __restrict__ leaf_parent, int* __restrict__ arrived, float* __restrict__ inner) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int node = leaf_parent[j];
    while (node >= 0) {
        __threadfence();                                  // this thread's box (written below) is visible before it announces itself
        if (atomicAdd(arrived + node, 1) == 0) return;    // the first to arrive leaves; the second finds both children complete
        __threadfence();
        const int2 c = children[node];
        const Fbox a = child_box(c.x, boxes, prims, inner), b = child_box(c.y, boxes, prims, inner);
        float* o = inner + 6 * (size_t)node;
        for (int k = 0; k < 3; ++k) { o[k] = fminf(a.lo[k], b.lo[k]); o[3 + k] = fmaxf(a.hi[k], b.hi[k]); }
        node = inner_parent[node];
    }
}

// One level of the collapse. work_in[w] = (wide node id, binary inner node); children that stay inner nodes go to work_out.
__global__ void k_collapse(const int2* __restrict__ work_in, int n_in, int2* __restrict__ work_out, int* __restrict__ counters /* [0] wide nodes, [1] work_out entries */,
                           int node_capacity, const float* __restrict__ boxes, const int* __restrict__ prims, const int2* __restrict__ children, const int2* __restrict__ range,
                           const float* __restrict__ inner, Node8* __restrict__ nodes) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_in) return;
    const int2 job = work_in[w];
    int kids[8]; int nk = 2;
    { const int2 c = children[job.y]; kids[0] = c.x; kids[1] = c.y; }
    auto size_of = [&](int ref) { if (ref < 0) return 1; const int2 r = range[ref]; return r.y - r.x + 1; };
    while (nk < 8) {   // open the child with the largest surface area among those that are not leaves yet
        int best = -1; float best_area = -1.0f;
        for (int i = 0; i < nk; ++i) {
            if (size_of(kids[i]) <= 4) continue;
            const float* b = inner + 6 * (size_t)kids[i];
            const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
            const float a = dx * dy + dx * dz + dy * dz;
            if (a > best_area) { best_area = a; best = i; }
        }
        if (best < 0) break;
        const int2 c = children[kids[best]];
        kids[best] = c.x; kids[nk++] = c.y;
    }
    if (job.x >= node_capacity) return;
    Node8 out;
    for (int k = 0; k < 6; ++k) for (int c = 0; c < 8; ++c) out.bounds[k][c] = (k & 1) ? -FLT_MAX : FLT_MAX;   // empty lanes: inverted finite box (bvh8.h)
    for (int c = 0; c < 8; ++c) { out.child[c] = 0; out.pad[c] = 0; }
    for (int i = 0; i < nk; ++i) {
        const Fbox b = child_box(kids[i], boxes, prims, inner);
        out.bounds[0][i] = b.lo[0]; out.bounds[1][i] = b.hi[0]; out.bounds[2][i] = b.lo[1]; out.bounds[3][i] = b.hi[1]; out.bounds[4][i] = b.lo[2]; out.bounds[5][i] = b.hi[2];
        const int sz = size_of(kids[i]);
        if (sz <= 4) {
            const int first = kids[i] < 0 ? ~kids[i] : range[kids[i]].x;
            out.child[i] = -(((first << 2) | (sz - 1)) + 1);
        } else {
            const int id = atomicAdd(counters, 1);
            out.child[i] = id + 1;
            work_out[atomicAdd(counters + 1, 1)] = make_int2(id, kids[i]);
        }
    }
    nodes[job.x] = out;
}

template <class T> struct Dev {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~Dev() { if (p) cudaFree(p); }
};

}  // namespace

#define BCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + " failed: " + cudaGetErrorString(e_); return false; } } while (0)

bool build_bvh8_gpu(const std::vector<Box3>& host_boxes, Bvh8& out, cudaStream_t stream, std::string& err) {
    const int n = (int)host_boxes.size();
    if (n < 5) { err = "build_bvh8_gpu: fewer than five primitives (such a shape is a single leaf)"; return false; }
    if (n >= (1 << 29)) { err = "build_bvh8_gpu: too many primitives for the 29-bit slot of a leaf code"; return false; }
    static_assert(sizeof(Box3) == 24, "Box3 is six packed floats");
    Dev<float> boxes, inner; Dev<unsigned> bounds, codes, codes2; Dev<int> prims, prims2, inner_parent, leaf_parent, arrived, counters; Dev<int2> children, range, work[2]; Dev<Node8> nodes; Dev<unsigned char> tmp;
    const int node_capacity = n / 2 + 16;
    BCU(boxes.alloc(6 * (size_t)n)); BCU(inner.alloc(6 * (size_t)n)); BCU(bounds.alloc(6)); BCU(codes.alloc(n)); BCU(codes2.alloc(n)); BCU(prims.alloc(n)); BCU(prims2.alloc(n));
    BCU(inner_parent.alloc(n)); BCU(leaf_parent.alloc(n)); BCU(arrived.alloc(n)); BCU(counters.alloc(2)); BCU(children.alloc(n)); BCU(range.alloc(n));
    BCU(work[0].alloc(n / 4 + 16)); BCU(work[1].alloc(n / 4 + 16)); BCU(nodes.alloc(node_capacity));
    BCU(cudaMemcpyAsync(boxes.p, host_boxes.data(), sizeof(Box3) * (size_t)n, cudaMemcpyHostToDevice, stream));
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    BCU(cudaMemcpyAsync(bounds.p, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    BCU(cudaMemsetAsync(arrived.p, 0, sizeof(int) * (size_t)n, stream));
    const int B = 256, G = (n + B - 1) / B;
    k_centroid_bounds<<<std::min(G, 1184), B, 0, stream>>>(boxes.p, n, bounds.p);
    k_morton<<<G, B, 0, stream>>>(boxes.p, n, bounds.p, codes.p, prims.p);
    size_t tmp_bytes = 0;
    BCU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes.p, codes2.p, prims.p, prims2.p, n, 0, 30, stream));
    BCU(tmp.alloc(tmp_bytes));
    BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, codes.p, codes2.p, prims.p, prims2.p, n, 0, 30, stream));
    k_radix_tree<<<G, B, 0, stream>>>(codes2.p, n, children.p, range.p, inner_parent.p, leaf_parent.p);
    k_fit<<<G, B, 0, stream>>>(n, boxes.p, prims2.p, children.p, inner_parent.p, leaf_parent.p, arrived.p, inner.p);
    // collapse, breadth first: the root is wide node 0 over binary node 0 (n > 4, so it is no leaf)
    const int2 root = make_int2(0, 0);
    BCU(cudaMemcpyAsync(work[0].p, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
    int n_nodes = 1, n_work = 1, depth = 0;
    for (int level = 0; n_work > 0; ++level) {
        const int c0[2] = {n_nodes, 0};
        BCU(cudaMemcpyAsync(counters.p, c0, sizeof(c0), cudaMemcpyHostToDevice, stream));
        k_collapse<<<(n_work + 127) / 128, 128, 0, stream>>>(work[level & 1].p, n_work, work[(level + 1) & 1].p, counters.p, node_capacity, boxes.p, prims2.p, children.p, range.p, inner.p, nodes.p);
        int c1[2];
        BCU(cudaMemcpyAsync(c1, counters.p, sizeof(c1), cudaMemcpyDeviceToHost, stream));
        BCU(cudaStreamSynchronize(stream));
        n_nodes = c1[0]; n_work = c1[1]; ++depth;
        if (n_nodes > node_capacity) { err = "build_bvh8_gpu: more wide nodes than the builder reserved (degenerate input); use the host builder"; return false; }
        if (level > 4096) { err = "build_bvh8_gpu: collapse does not terminate"; return false; }
    }
    BCU(cudaGetLastError());
    out.nodes.resize((size_t)n_nodes);
    out.order.resize((size_t)n);
    BCU(cudaMemcpyAsync(out.nodes.data(), nodes.p, sizeof(Node8) * (size_t)n_nodes, cudaMemcpyDeviceToHost, stream));
    BCU(cudaMemcpyAsync(out.order.data(), prims2.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, stream));
    BCU(cudaStreamSynchronize(stream));
    out.max_depth = depth;
    return true;
}

// ---- on-disk cache ---------------------------------------------------------------------------------------------------------------
// One file per tree, named by the hash of what the tree is a function of (the primitive boxes, the leaf size and the builder), so an
// entry can never be stale: <dir>/bvh8_<hash>.bin = header | nodes | order. The reference keys `_bvh_<shape name>.bin` by a separate
// hash registry (CacheManager::checkAndUpdate); content addressing needs no registry and is safe when several processes (one per GPU)
// load the same scene at once -- files are written under a temporary name and renamed.
namespace {
struct CacheHeader { char magic[8]; uint32_t version, builder, n_prims, n_nodes, max_depth, pad; uint64_t hash; };
const char kMagic[8] = {'I', 'G', 'B', '2', '0', '0', 'B', '8'};
std::string cache_path(const std::string& dir, uint64_t hash) { char name[64]; snprintf(name, sizeof(name), "/bvh8_%016llx.bin", (unsigned long long)hash); return dir + name; }
}  // namespace

uint64_t bvh_cache_hash(const std::vector<Box3>& boxes, int max_leaf, int builder) {
    uint64_t h = 1469598103934665603ull;   // FNV-1a over 8-byte words (the boxes are 24 bytes each), then the parameters
    auto mix = [&](uint64_t w) { h ^= w; h *= 1099511628211ull; };
    const size_t bytes = boxes.size() * sizeof(Box3);
    const unsigned char* p = reinterpret_cast<const unsigned char*>(boxes.data());
    for (size_t i = 0; i + 8 <= bytes; i += 8) { uint64_t w; memcpy(&w, p + i, 8); mix(w); }
    mix((uint64_t)boxes.size()); mix((uint64_t)max_leaf); mix((uint64_t)builder);
    return h;
}

bool bvh_cache_load(const std::string& dir, uint64_t hash, size_t n_prims, int builder, Bvh8& out) {
    FILE* f = fopen(cache_path(dir, hash).c_str(), "rb");
    if (!f) return false;
    CacheHeader hd;
    bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, kMagic, 8) == 0 && hd.version == 1 && hd.hash == hash && hd.n_prims == n_prims && hd.builder == (uint32_t)builder &&
              hd.n_nodes >= 1 && hd.n_nodes <= n_prims + 16;
    if (ok) {
        out.nodes.resize(hd.n_nodes); out.order.resize(n_prims); out.max_depth = (int)hd.max_depth;
        ok = fread(out.nodes.data(), sizeof(Node8), hd.n_nodes, f) == hd.n_nodes && fread(out.order.data(), sizeof(int32_t), n_prims, f) == n_prims && fgetc(f) == EOF;
    }
    fclose(f);
    if (ok) {   // a file is trusted only as far as it cannot make the traversal leave its arrays
        for (const Node8& nd : out.nodes)
            for (int k = 0; k < 8 && ok; ++k) {
                const int c = nd.child[k];
                if (c > 0) ok = (uint32_t)c <= hd.n_nodes;
                else if (c < 0) { const int r = -c - 1; ok = (size_t)(r >> 2) + (size_t)(r & 3) + 1 <= n_prims; }
            }
        for (size_t i = 0; i < n_prims && ok; ++i) ok = out.order[i] >= 0 && (size_t)out.order[i] < n_prims;
    }
    if (!ok) { out.nodes.clear(); out.order.clear(); out.max_depth = 0; }
    return ok;
}

bool bvh_cache_store(const std::string& dir, uint64_t hash, int builder, const Bvh8& bvh) {
    const std::string path = cache_path(dir, hash);
    char suffix[48]; snprintf(suffix, sizeof(suffix), ".tmp.%ld.%p", (long)getpid(), (const void*)&bvh);
    const std::string tmp = path + suffix;
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    CacheHeader hd; memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, kMagic, 8); hd.version = 1; hd.builder = (uint32_t)builder; hd.n_prims = (uint32_t)bvh.order.size(); hd.n_nodes = (uint32_t)bvh.nodes.size(); hd.max_depth = (uint32_t)bvh.max_depth; hd.hash = hash;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 && fwrite(bvh.nodes.data(), sizeof(Node8), bvh.nodes.size(), f) == bvh.nodes.size() &&
              fwrite(bvh.order.data(), sizeof(int32_t), bvh.order.size(), f) == bvh.order.size();
    ok = fclose(f) == 0 && ok;
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;
    if (!ok) remove(tmp.c_str());
    return ok;
}

}  // namespace igb
