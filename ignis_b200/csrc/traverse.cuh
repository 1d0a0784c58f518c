// Two-level BVH8 traversal as a resumable per-lane state machine (one `step` = one node, entity or leaf visit).
//
// Reference behaviour implemented (paths relative to the reference's src/artic):
//   traversal/mapping_cpu.art:421-518 (top level), :282-412 (bottom level), traversal/mapping_gpu.art:67-219,
//   traversal/intersection.art:74-106 (Moeller-Trumbore), :223-256 (slabs), traversal/ray.art:27-59,
//   shapes/trimesh.art:124-144, shapes/sphere.art:108-147.
//
// B200 design:
//  * One loop walks both levels: an entity leaf of the top-level tree switches the ray to the entity's local space
//    (ray.art:53-59) and pushes a sentinel that switches back when popped, so lanes in either level execute the
//    same node and triangle code. Entities whose local matrix is exactly the identity keep the world ray (the
//    transform is the exact map x -> x + 0), shapes of <= 4 triangles have no inner node at all.
//  * The state machine is resumable so that the persistent trace phase (wavefront.cuh) can refill finished lanes
//    of a warp with new rays while the others are still walking ("persistent threads", Aila & Laine 2009).
//  * The traversal stack lives in shared memory, interleaved by thread (entry k of thread t at [k * stride + t]),
//    so a push/pop is one conflict-free wavefront whatever the lanes' stack depths are; entries beyond SMEM_STACK
//    spill to a per-thread local array. Nodes, triangles and entity leaves are read from a shared-memory copy
//    staged by TMA bulk copies when the scene (or its first part) fits, else from global memory.
//  * Pruning against the running closest hit is CONSERVATIVE (DESIGN.md "Ties and conservative culling"): slab
//    distances are widened by a bound of their rounding error (|inv_org| * 2^-20 per axis) and compared with
//    t_closest * (1 + 2^-16), so a box test can never hide a primitive whose Moeller-Trumbore distance is <= the
//    current one. The result is a pure function of the ray -- min over all primitives ordered by (t, entity,
//    primitive) -- and equals the oracle's brute-force answer whatever the tree looks like. The entity-box test
//    of the reference (mapping_cpu.art:480, intersection.art:247-256) is part of the semantics and kept bit for bit.
#pragma once

#include "types.cuh"

namespace igb {

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


// ---- stack -----------------------------------------------------------------------------------------------------
constexpr int SMEM_STACK  = 16;                       // entries per thread held in shared memory
constexpr int LOCAL_STACK = 80;                       // overflow entries per thread (local memory)
constexpr int STACK_SIZE  = SMEM_STACK + LOCAL_STACK;
constexpr int   SENTINEL_RESTORE   = (int)0x80000000;   // pop: back to world space, reload the ray
constexpr int   SENTINEL_RESTORE_T = (int)0x80000001;   // pop: back to world space, only the origin changed (translated instance)
constexpr int   SENTINEL_KEEP      = (int)0x80000002;   // pop: back to the top level, ray unchanged (identity instance); the largest sentinel
constexpr float CULL_SLACK = 9.5367431640625e-07f;    // 2^-20
constexpr float CULL_TMAX  = 1.0000152587890625f;     // 1 + 2^-16

struct Stack {
    uint2* s;       // this thread's column in shared memory
    int    stride;  // threads per block
    uint2* l;       // this thread's local overflow array
    __device__ __forceinline__ void push(int& sp, int node, float t) {
        const uint2 e = make_uint2((uint32_t)node, __float_as_uint(t));
        if (sp < SMEM_STACK) s[sp * stride] = e; else l[sp - SMEM_STACK] = e;
        ++sp;
    }
    __device__ __forceinline__ uint2 pop(int& sp) {
        --sp;
        return sp < SMEM_STACK ? s[sp * stride] : l[sp - SMEM_STACK];
    }
    __device__ __forceinline__ void set(int slot, uint2 e) {
        if (slot < SMEM_STACK) s[slot * stride] = e; else l[slot - SMEM_STACK] = e;
    }
};

// Scene arrays (or their first part) staged in shared memory; indices beyond the staged count read global memory.
// The staged copies are PADDED: a node (256 B) every 272 bytes, an entity leaf (128 B) every 144 bytes. Lanes of a warp read the same row of
// DIFFERENT nodes (e.g. the near x planes, a float4 each); with a stride that is a multiple of 128 bytes all of those land in the same four
// banks and the load is replayed once per lane (ncu, unpadded: 55 M bank conflicts and 116 M shared-load wavefronts per launch, the LSU shared
// pipe 59 % busy; padded: 19 M / 80 M / 44 % -- profiles/r6_trace_experiments.txt). One extra float4 per element rotates the bank group with
// the index. Triangles (48 B) are spread already. The kernel's time did not move: it is not bound by the shared-memory pipe.
constexpr int STAGED_NODE_F4 = 17, STAGED_LEAF_F4 = 9;
constexpr int STAGED_NODE_BYTES = STAGED_NODE_F4 * 16, STAGED_LEAF_BYTES = STAGED_LEAF_F4 * 16;
struct Staged {
    const float4* nodes;    int n_nodes;
    const float4* tris;     int n_tris;
    const float4* ent_leaf; int n_ent;
};
// WHERE: 0 = per index (a staged prefix, the rest in global memory: generic loads), 1 = the whole scene is staged (shared-memory
// loads, no range checks), 2 = nothing is staged (global loads). The host stages all or nothing (api.cu configure_kernels), so 1 and
// 2 are what actually runs; 0 remains for the "stage_partial" option and the kernels that are not specialised.
template <int WHERE = 0>
__device__ __forceinline__ const float4* node_ptr(const DevScene& sc, const Staged& sg, int node) {
    if (WHERE == 1) return sg.nodes + node * STAGED_NODE_F4;
    if (WHERE == 2) return sc.nodes + (size_t)node * 16;
    return node < sg.n_nodes ? sg.nodes + node * STAGED_NODE_F4 : sc.nodes + (size_t)node * 16;
}
template <int WHERE = 0>
__device__ __forceinline__ const float4* tri_ptr(const DevScene& sc, const Staged& sg, int slot) {
    if (WHERE == 1) return sg.tris + slot * 3;
    if (WHERE == 2) return sc.tris + (size_t)slot * 3;
    return slot < sg.n_tris ? sg.tris + slot * 3 : sc.tris + (size_t)slot * 3;
}
template <int WHERE = 0>
__device__ __forceinline__ const float4* leaf_ptr(const DevScene& sc, const Staged& sg, int slot) {
    if (WHERE == 1) return sg.ent_leaf + slot * STAGED_LEAF_F4;
    if (WHERE == 2) return sc.ent_leaf + (size_t)slot * 8;
    return slot < sg.n_ent ? sg.ent_leaf + slot * STAGED_LEAF_F4 : sc.ent_leaf + (size_t)slot * 8;
}

// Ray state bits
constexpr uint32_t TB_ANY = 1u << 8, TB_NEGZERO = 1u << 9, TB_SX = 1u << 10, TB_SY = 1u << 11, TB_SZ = 1u << 12, TB_DONE = 1u << 13,
                   TB_STALE_T = 1u << 15,  // ... only its origin (the instance was a pure translation)
                   TB_STALE = 1u << 14;   // an instance was left: the world-space ray must be restored before the next visit (lazily: often there is none)

// One slab test of child lane K of a group of four; the entry is written to the stack slot above the ones pushed so far
// and only kept (n advances) if the child is hit, so there is no branch per child. `nearer` uses "not >=" so that the
// first hit child replaces the NaN the running minimum starts from.
#define IGB_CHILD(K)                                                                                                     \
    {                                                                                                                    \
        const float tn = fmaxf(fmaxf(fma_(idir.x, nx.K, ilo.x), fma_(idir.y, ny.K, ilo.y)), fmaxf(fma_(idir.z, nz.K, ilo.z), tmin));   \
        const float tf = fminf(fminf(fma_(idir.x, fx.K, ihi.x), fma_(idir.y, fy.K, ihi.y)), fminf(fma_(idir.z, fz.K, ihi.z), tcull));  \
        const bool hit_ = tn <= tf;                                                                                      \
        IGB_STORE(ch.K, tn)                                                                                              \
        const bool nearer = hit_ && !(tn >= best_t);                                                                     \
        best_t = nearer ? tn : best_t; best = nearer ? ch.K : best; best_slot = nearer ? n : best_slot;                  \
        n += hit_ ? 1 : 0;                                                                                               \
    }

struct Traversal {
    V3 org, dir;              // ray in the current space (world, or local to `ent`)
    float tmin, tmax;         // the ray's own interval (never shrunk: the triangle test uses it, see header)
    V3 idir, ilo, ihi;        // 1/dir, -(org/dir) minus / plus its error bound (conservative slabs)
    float tcull;              // hit.t * CULL_TMAX
    HitR hit;
    int cur, ent, sp;         // node code (0: pop next); entity whose bottom-level tree is walked (-1: top level)
                              // (merged-tree walk: `ent` = the entity-leaf slot whose local ray is cached in lorg / ldir, -1: none)
    V3 lorg, ldir;            // merged-tree walk only: the ray in the local space of entity slot `ent`
    uint32_t bits;            // ray type flags (ray.art:51) | TB_*
#ifdef IGB_STEP_STATS
    int n_node, n_leaf, n_ent;   // diagnostics build only: visits of each kind
#endif

    __device__ __forceinline__ void set_ray(V3 o, V3 d) {
        dir = d;
        idir = v3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));                 // traversal/ray.art:27-39
        bits = (bits & ~(TB_SX | TB_SY | TB_SZ)) | (idir.x < 0 ? TB_SX : 0u) | (idir.y < 0 ? TB_SY : 0u) | (idir.z < 0 ? TB_SZ : 0u);
        set_origin(o);
    }
    // the part of set_ray that depends on the origin: enough when the direction is unchanged (translated instances)
    __device__ __forceinline__ void set_origin(V3 o) {
        org = o;
        const V3 iorg = neg(o * idir);
        const float inf = __int_as_float(0x7f800000);
        const float ax = fabsf(iorg.x), ay = fabsf(iorg.y), az = fabsf(iorg.z);
        // |iorg| = inf: org * flt_max overflowed (axis-parallel ray) -> no bound from that axis
        ilo = v3(ax == inf ? -inf : iorg.x - ax * CULL_SLACK, ay == inf ? -inf : iorg.y - ay * CULL_SLACK, az == inf ? -inf : iorg.z - az * CULL_SLACK);
        ihi = v3(ax == inf ? inf : iorg.x + ax * CULL_SLACK, ay == inf ? inf : iorg.y + ay * CULL_SLACK, az == inf ? inf : iorg.z + az * CULL_SLACK);
    }

    // po / pd: the ray record (org.xyz,tmin / dir.xyz,tmax). Returns false if there is nothing to traverse.
    __device__ __forceinline__ bool begin(const DevScene& sc, const float4* po, const float4* pd, uint32_t ray_flags, bool any_hit) {
        const float4 o = *po, d = *pd;
        tmin = o.w; tmax = d.w;
        hit.t = tmax; hit.u = 0; hit.v = 0; hit.prim = -1; hit.ent = -1;
        tcull = hit.t * CULL_TMAX;
        cur = 1; ent = -1; sp = 0;
#ifdef IGB_STEP_STATS
        n_node = n_leaf = n_ent = 0;
#endif
        const bool nz = ((__float_as_uint(o.x) == 0x80000000u) | (__float_as_uint(o.y) == 0x80000000u) | (__float_as_uint(o.z) == 0x80000000u) |
                         (__float_as_uint(d.x) == 0x80000000u) | (__float_as_uint(d.y) == 0x80000000u) | (__float_as_uint(d.z) == 0x80000000u));
        bits = (ray_flags & RAY_TYPE_MASK) | (any_hit ? TB_ANY : 0u) | (nz ? TB_NEGZERO : 0u);
        if (sc.n_ent == 0) return false;
        set_ray(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z));
        return true;
    }

    __device__ __forceinline__ void accept(float t, float u, float v, int prim, int e) {
        hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; hit.ent = e; tcull = t * CULL_TMAX;
        if (bits & TB_ANY) bits |= TB_DONE;
    }

    // ---- P: pop the next node. Returns true when the stack is empty (traversal finished).
    __device__ __forceinline__ bool pop_step(Stack& st, const float4* po, const float4* pd) {
        for (;;) {
            if (sp == 0) return true;
            const uint2 e = st.pop(sp);
            cur = (int)e.x;
            if (cur > SENTINEL_KEEP) {
                if (!(__uint_as_float(e.y) <= tcull)) continue;
                // leaving a transformed instance costs three reciprocals: pay only if something is still to be visited in world space
                if (bits & (TB_STALE | TB_STALE_T)) {
                    const float4 o = *po;
                    if (bits & TB_STALE) { const float4 d = *pd; set_ray(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z)); }
                    else set_origin(v3(o.x, o.y, o.z));
                    bits &= ~(TB_STALE | TB_STALE_T);
                }
                return false;
            }
            ent = -1;
            if (cur == SENTINEL_RESTORE) bits |= TB_STALE;
            else if (cur == SENTINEL_RESTORE_T) bits |= TB_STALE_T;
        }
    }

    // ---- N for nodes in GLOBAL memory (WHERE = 2, large scenes): the same eight slab tests, fed by 256-bit loads. A row of the node (one
    // plane of all eight children) is 32 bytes = one sector; the float4 walk below reads it with two instructions, i.e. two passes through the
    // L1 data pipe for the same sector -- and that pipe is what ncu shows saturated on unstaged scenes (l1tex data-pipe wavefronts 75 % of peak,
    // long scoreboard the top stall, profiles/r5h_room_k_turn_trace_full.csv). sm_100 has ld.global.v8.f32 (LDG.E.256): seven loads per node
    // instead of fourteen. Evaluated plane by plane (near / far of one axis at a time) so that 16 loaded values are live, not 48.
    __device__ __forceinline__ static void ldg8(const float4* p, float (&v)[8]) {
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
    }
    __device__ __forceinline__ void node_step_global8(const DevScene& sc, Stack& st) {
        const float4* N = sc.nodes + (size_t)(cur - 1) * 16;
        const int ox = (bits & TB_SX) ? 2 : 0, oy = (bits & TB_SY) ? 2 : 0, oz = (bits & TB_SZ) ? 2 : 0;
        float tn[8], tf[8], a[8], b[8], chf[8];
        ldg8(N + ox, a); ldg8(N + 2 - ox, b); ldg8(N + 12, chf);
#pragma unroll
        for (int k = 0; k < 8; ++k) { tn[k] = fma_(idir.x, a[k], ilo.x); tf[k] = fma_(idir.x, b[k], ihi.x); }
        ldg8(N + 4 + oy, a); ldg8(N + 6 - oy, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) { tn[k] = fmaxf(tn[k], fma_(idir.y, a[k], ilo.y)); tf[k] = fminf(tf[k], fma_(idir.y, b[k], ihi.y)); }
        ldg8(N + 8 + oz, a); ldg8(N + 10 - oz, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) { tn[k] = fmaxf(tn[k], fmaxf(fma_(idir.z, a[k], ilo.z), tmin)); tf[k] = fminf(tf[k], fminf(fma_(idir.z, b[k], ihi.z), tcull)); }
        int n = 0, best = 0, best_slot = 0;
        float best_t = __int_as_float(0x7fc00000);
        if (sp + 8 <= SMEM_STACK) {
            uint2* base = st.s + sp * st.stride;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = __float_as_int(chf[k]);
                const bool hit_ = tn[k] <= tf[k];
                base[n * st.stride] = make_uint2((uint32_t)c, __float_as_uint(tn[k]));
                const bool nearer = hit_ && !(tn[k] >= best_t);
                best_t = nearer ? tn[k] : best_t; best = nearer ? c : best; best_slot = nearer ? n : best_slot;
                n += hit_ ? 1 : 0;
            }
            if (n == 0) { cur = 0; return; }
            --n;
            if (best_slot != n) base[best_slot * st.stride] = base[n * st.stride];
            sp += n; cur = best;
        } else {
            const int sp0 = sp;
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const int c = __float_as_int(chf[k]);
                const bool hit_ = tn[k] <= tf[k];
                if (hit_ && sp < STACK_SIZE) st.push(sp, c, tn[k]);
                const bool nearer = hit_ && !(tn[k] >= best_t);
                best_t = nearer ? tn[k] : best_t; best = nearer ? c : best; best_slot = nearer ? n : best_slot;
                n += hit_ ? 1 : 0;
            }
            if (sp == sp0) { cur = 0; return; }
            const uint2 top = st.pop(sp);
            if (sp0 + best_slot != sp) st.set(sp0 + best_slot, top);
            cur = best;
        }
    }

    // ---- N: inner node. Eight branch-free slab tests; the nearest hit child becomes `cur`, the others stay pushed.
    template <int WHERE = 0>
    __device__ __forceinline__ void node_step(const DevScene& sc, const Staged& sg, Stack& st) {
#ifdef IGB_STEP_STATS
        ++n_node;
#endif
        if (WHERE == 2) { node_step_global8(sc, st); return; }
        const float4* N = node_ptr<WHERE>(sc, sg, cur - 1);
        // rows of the node: lo_x hi_x lo_y hi_y lo_z hi_z, two float4 each; the octant picks near / far by address
        const int ox = (bits & TB_SX) ? 2 : 0, oy = (bits & TB_SY) ? 2 : 0, oz = (bits & TB_SZ) ? 2 : 0;
        int n = 0, best = 0, best_slot = 0;
        float best_t = __int_as_float(0x7fc00000);
        if (sp + 8 <= SMEM_STACK) {
            uint2* base = st.s + sp * st.stride;
#define IGB_STORE(C, T) base[n * st.stride] = make_uint2((uint32_t)(C), __float_as_uint(T));
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int4 ch = *reinterpret_cast<const int4*>(N + 12 + g);
                if (g == 1 && ch.x == 0) break;   // children are packed to the front
                const float4 nx = N[ox + g], fx = N[2 - ox + g], ny = N[4 + oy + g], fy = N[6 - oy + g], nz = N[8 + oz + g], fz = N[10 - oz + g];
                IGB_CHILD(x) IGB_CHILD(y) IGB_CHILD(z) IGB_CHILD(w)
            }
#undef IGB_STORE
            if (n == 0) { cur = 0; return; }
            --n;
            if (best_slot != n) base[best_slot * st.stride] = base[n * st.stride];
            sp += n; cur = best;
        } else {
            // deep stack: same tests, entries go through the overflow-aware push (rare)
            const int sp0 = sp;
#define IGB_STORE(C, T) if (hit_ && sp < STACK_SIZE) st.push(sp, (C), (T));
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                const int4 ch = *reinterpret_cast<const int4*>(N + 12 + g);
                if (g == 1 && ch.x == 0) break;
                const float4 nx = N[ox + g], fx = N[2 - ox + g], ny = N[4 + oy + g], fy = N[6 - oy + g], nz = N[8 + oz + g], fz = N[10 - oz + g];
                IGB_CHILD(x) IGB_CHILD(y) IGB_CHILD(z) IGB_CHILD(w)
            }
#undef IGB_STORE
            if (sp == sp0) { cur = 0; return; }
            const uint2 top = st.pop(sp);
            if (sp0 + best_slot != sp) st.set(sp0 + best_slot, top);
            cur = best;
        }
    }

    // ---- E: top-level leaf = one entity (traversal/mapping_cpu.art:470-500)
    template <int WHERE = 0>
    __device__ __forceinline__ void entity_step(const DevScene& sc, const Staged& sg, Stack& st) {
#ifdef IGB_STEP_STATS
        ++n_ent;
#endif
        const float4* L = leaf_ptr<WHERE>(sc, sg, (-cur - 1) >> 2);
        const float4 l0 = L[0], l1 = L[1];
        cur = 0;
        const uint32_t eflags = __float_as_uint(l0.w);
        if ((bits & RAY_TYPE_MASK) != ((bits & eflags) & RAY_TYPE_MASK)) return;                      // ray.art:51
        {   // intersect_ray_box_single_section, intersection.art:247-256, against the ray's own tmax (exact reference form)
            const V3 iorg = neg(org * idir);
            const float t0x = fma_(idir.x, l0.x, iorg.x), t1x = fma_(idir.x, l1.x, iorg.x);
            const float t0y = fma_(idir.y, l0.y, iorg.y), t1y = fma_(idir.y, l1.y, iorg.y);
            const float t0z = fma_(idir.z, l0.z, iorg.z), t1z = fma_(idir.z, l1.z, iorg.z);
            const float en = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
            const float ex = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
            if (!((en <= ex) & (ex >= 0))) return;
        }
        const float4 l5 = L[5];
        const int kind = __float_as_int(l1.w);    // bit 0: analytic sphere, bit 1: identity local matrix, bit 2: translation only
        const int e = __float_as_int(l5.x);
        if (kind & 1) {
            const float4 r0 = L[2], r1 = L[3], r2 = L[4], s = L[6];
            float t, u, v;
            if (intersect_sphere(v3(s.x, s.y, s.z), s.w, xform_point(r0, r1, r2, org), xform_dir(r0, r1, r2, dir), tmin, tmax, t, u, v) && better(t, e, 0, hit))
                accept(t, u, v, 0, e);
            return;
        }
        ent = e;
        cur = __float_as_int(l5.y);               // root of the shape's tree (a leaf code for shapes of <= 4 triangles)
        if ((kind & 2) && !(bits & TB_NEGZERO)) {
            if (cur < 0) {
                // untransformed shape of <= 4 triangles (walls, area lights): test them right here -- no sentinel to push and
                // pop, no separate leaf visit, and one visit kind less for the warp to diverge over
                leaf_step<WHERE>(sc, sg);
                ent = -1;
                return;
            }
            st.push(sp, SENTINEL_KEEP, -1.0f);
            return;
        }
        const float4 r0 = L[2], r1 = L[3], r2 = L[4];
        if ((kind & 4) && !(bits & TB_NEGZERO)) {
            // pure translation: the transformed direction is the direction (exactly, as no component is -0), so only the origin moves
            set_origin(xform_point(r0, r1, r2, org));
            st.push(sp, SENTINEL_RESTORE_T, -1.0f);
            return;
        }
        set_ray(xform_point(r0, r1, r2, org), xform_dir(r0, r1, r2, dir));                            // ray.art:53-59
        st.push(sp, SENTINEL_RESTORE, -1.0f);
    }

    // ---- L: bottom-level leaf = up to four triangles (shapes/trimesh.art:124-144)
    template <int WHERE = 0>
    __device__ __forceinline__ void leaf_step(const DevScene& sc, const Staged& sg) {
#ifdef IGB_STEP_STATS
        ++n_leaf;
#endif
        const int r = -cur - 1;
        const int first = r >> 2, cnt = (r & 3) + 1;
        cur = 0;
        if (WHERE == 1 || (WHERE == 0 && first + cnt <= sg.n_tris)) {   // leaf staged in shared memory
            for (int j = 0; j < cnt; ++j) {
                const float4* T = sg.tris + (first + j) * 3;
                const float4 a = T[0], b = T[1], c = T[2];
                float t, u, v;
                if (intersect_tri(org, dir, tmin, tmax, a, b, c, t, u, v)) {
                    const int prim = __ldg(sc.tri_prim + first + j);
                    if (better(t, ent, prim, hit)) accept(t, u, v, prim, ent);
                }
            }
            return;
        }
        // geometry read through L2: one triangle of look-ahead, so that the next triangle's three loads are in flight while this
        // one is tested (else a leaf costs up to four dependent round trips; synthetic_room: -6 % trace time)
        const float4* T = tri_ptr<WHERE>(sc, sg, first);
        float4 a = T[0], b = T[1], c = T[2];
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            float4 na = a, nb = b, nc = c;
            if (j + 1 < cnt) { const float4* Tn = tri_ptr<WHERE>(sc, sg, first + j + 1); na = Tn[0]; nb = Tn[1]; nc = Tn[2]; }
            float t, u, v;
            if (intersect_tri(org, dir, tmin, tmax, a, b, c, t, u, v)) {
                const int prim = __ldg(sc.tri_prim + first + j);
                if (better(t, ent, prim, hit)) accept(t, u, v, prim, ent);
            }
            a = na; b = nb; c = nc;
        }
    }

    // ================= merged-tree walk (small scenes; DevScene::flat_nodes) =====================================================
    // The instances' trees hang directly off the top-level tree and their node boxes are refitted in world space, so the walk never
    // leaves world space: no entity visit, no sentinel, no ray set-up (three reciprocals) on entering or leaving an instance -- two visit
    // kinds instead of three for the warp to diverge over. What the reference does per entity (traversal/mapping_cpu.art:479-494) is
    // kept to the letter, only moved: the ray is transformed into the entity's space when a leaf of that entity is reached (cached
    // while consecutive leaves belong to the same entity) and the triangle test runs on that local ray exactly as before, so t / u / v
    // are the same bits; the entity's visibility and box tests are properties of (ray, entity) and are evaluated when a candidate hit
    // in that entity turns up. A leaf code carries the entity: r = -code - 1 = (entity-leaf slot << 22) | (first slot << 2) | (count - 1).
    __device__ __forceinline__ bool pop_flat(Stack& st) {
        for (;;) {
            if (sp == 0) return true;
            const uint2 e = st.pop(sp);
            if (__uint_as_float(e.y) <= tcull) { cur = (int)e.x; return false; }
        }
    }
    template <int WHERE>
    __device__ __forceinline__ bool entity_admits(const DevScene& sc, const Staged& sg, int slot) {
        const float4* L = leaf_ptr<WHERE>(sc, sg, slot);
        const float4 l0 = L[0], l1 = L[1];
        const uint32_t eflags = __float_as_uint(l0.w);
        if ((bits & RAY_TYPE_MASK) != ((bits & eflags) & RAY_TYPE_MASK)) return false;                       // ray.art:51
        const V3 iorg = neg(org * idir);                                                                       // intersection.art:247-256, exact reference form
        const float t0x = fma_(idir.x, l0.x, iorg.x), t1x = fma_(idir.x, l1.x, iorg.x);
        const float t0y = fma_(idir.y, l0.y, iorg.y), t1y = fma_(idir.y, l1.y, iorg.y);
        const float t0z = fma_(idir.z, l0.z, iorg.z), t1z = fma_(idir.z, l1.z, iorg.z);
        const float en = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
        const float ex = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
        return (en <= ex) & (ex >= 0);
    }
    template <int WHERE>
    __device__ __forceinline__ void flat_leaf_step(const DevScene& sc, const Staged& sg) {
        const int r = -cur - 1;
        const int cnt = (r & 3) + 1, first = (r >> 2) & 0xFFFFF, slot = r >> 22;
        cur = 0;
        const float4* L = leaf_ptr<WHERE>(sc, sg, slot);
        if (slot != ent) {
            const int kind = __float_as_int(L[1].w);
            if ((kind & 2) && !(bits & TB_NEGZERO)) { lorg = org; ldir = dir; }                              // bit-exact identity: x -> x + 0
            else { const float4 r0 = L[2], r1 = L[3], r2 = L[4]; lorg = xform_point(r0, r1, r2, org); ldir = xform_dir(r0, r1, r2, dir); }   // ray.art:53-59
            ent = slot;
        }
        if (WHERE == 1) {
            for (int j = 0; j < cnt; ++j) {
                const float4* T = tri_ptr<WHERE>(sc, sg, first + j);
                const float4 a = T[0], b = T[1], c = T[2];
                float t, u, v;
                if (intersect_tri(lorg, ldir, tmin, tmax, a, b, c, t, u, v)) {
                    const int prim = __ldg(sc.tri_prim + first + j);
                    const int e = __float_as_int(L[5].x);
                    if (better(t, e, prim, hit) && entity_admits<WHERE>(sc, sg, slot)) accept(t, u, v, prim, e);
                }
            }
            return;
        }
        // geometry read through L2: one triangle of look-ahead, as in leaf_step
        const float4* T = tri_ptr<WHERE>(sc, sg, first);
        float4 a = T[0], b = T[1], c = T[2];
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            float4 na = a, nb = b, nc = c;
            if (j + 1 < cnt) { const float4* Tn = tri_ptr<WHERE>(sc, sg, first + j + 1); na = Tn[0]; nb = Tn[1]; nc = Tn[2]; }
            float t, u, v;
            if (intersect_tri(lorg, ldir, tmin, tmax, a, b, c, t, u, v)) {
                const int prim = __ldg(sc.tri_prim + first + j);
                const int e = __float_as_int(L[5].x);
                if (better(t, e, prim, hit) && entity_admits<WHERE>(sc, sg, slot)) accept(t, u, v, prim, e);
            }
            a = na; b = nb; c = nc;
        }
    }
    template <int WHERE>
    __device__ __forceinline__ bool turn_vote_flat(const DevScene& sc, const Staged& sg, Stack& st, bool active, int vote) {
        bool fin = false;
        if (active && cur == 0) fin = pop_flat(st);
        const bool live = active && !fin;
        const bool wN = live && cur > 0, wL = live && cur < 0;
        const int nN = __popc(__ballot_sync(0xffffffffu, wN)), nL = __popc(__ballot_sync(0xffffffffu, wL));
        const int mx = max(nN, nL);
        const int need = vote >= 2 ? max(1, (mx + 1) >> 1) : max(1, mx);
        if (nN >= need) { if (wN) node_step<WHERE>(sc, sg, st); }
        if (nL >= need) { if (cur < 0 && live) flat_leaf_step<WHERE>(sc, sg); }
        return fin || (live && (bits & TB_DONE) != 0);
    }

    // One turn of the pipeline P -> N -> E -> L: a lane performs every visit its state allows, in that order, so that
    // a fresh ray does root node, entity and triangle leaf in one turn and lanes of a warp stay aligned.
    // Returns true when the traversal is finished (hit holds the result).
    template <int WHERE = 0>
    __device__ __forceinline__ bool turn(const DevScene& sc, const Staged& sg, Stack& st, const float4* po, const float4* pd) {
        if (cur == 0 && pop_step(st, po, pd)) return true;
        if (cur > 0) node_step<WHERE>(sc, sg, st);
        if (cur < 0 && ent < 0) entity_step<WHERE>(sc, sg, st);
        if (cur < 0 && ent >= 0) leaf_step<WHERE>(sc, sg);
        return (bits & TB_DONE) != 0;
    }

    // The same pipeline scheduled by warp vote: lanes pop, then ONE visit kind (vote 1: the most wanted; vote 2: every
    // kind wanted by at least half as many lanes as the most wanted) runs, lanes waiting for another kind keep their
    // state. The kinds then run 2-3x wider than when all three sections execute back to back for whoever needs them.
    // Must be called by all 32 lanes; `active` false lanes only vote. Returns true when this lane's traversal finished.
    template <int WHERE = 0>
    __device__ __forceinline__ bool turn_vote(const DevScene& sc, const Staged& sg, Stack& st, const float4* po, const float4* pd, bool active, int vote) {
        bool fin = false;
        if (active && cur == 0) fin = pop_step(st, po, pd);
        const bool live = active && !fin;
        const bool wN = live && cur > 0, wE = live && cur < 0 && ent < 0, wL = live && cur < 0 && ent >= 0;
        const int nN = __popc(__ballot_sync(0xffffffffu, wN)), nE = __popc(__ballot_sync(0xffffffffu, wE)), nL = __popc(__ballot_sync(0xffffffffu, wL));
        const int mx = max(nN, max(nE, nL));
        const int need = vote >= 2 ? max(1, (mx + 1) >> 1) : max(1, mx);
        if (nN >= need) { if (wN) node_step<WHERE>(sc, sg, st); }
        if (nE >= need) { if (cur < 0 && ent < 0 && live) entity_step<WHERE>(sc, sg, st); }
        if (nL >= need) { if (cur < 0 && ent >= 0 && live) leaf_step<WHERE>(sc, sg); }
        return fin || (live && (bits & TB_DONE) != 0);
    }
};
#undef IGB_CHILD

// ---- wide traversal: EIGHT lanes walk ONE ray --------------------------------------------------------------------
// Used for small queues (the deep-path tail of an iteration, wavefront.cuh): there the trace phase lasts exactly as long
// as the slowest single ray, and a scalar lane needs ~600 dependent instructions per loop turn (eight slab tests, up to
// four triangle tests). With a BVH8, lane j of a group of eight tests child j of the node / triangle j of the leaf, the
// nearest child is found with three shuffles, and the critical path per visit shrinks by the same factor.
// The closest hit is a pure function of the ray (see the header of this file), so the visiting order -- which differs
// from the scalar walk -- cannot change the result; tests/test_gpu_parity.py checks both walks against the oracle.
// All state is replicated in the eight lanes; control flow is uniform within a group and divergent between groups,
// hence every warp primitive below uses the group's 8-lane mask.
struct WideStack {
    uint2* base;   // shared memory: the stack columns of the group's first thread (16 rows x 8 columns = 128 entries)
    int stride;
    __device__ __forceinline__ uint2& at(int k) const { return base[(k >> 3) * stride + (k & 7)]; }
};
constexpr int WIDE_STACK = SMEM_STACK * 8;

template <int WHERE = 0>
__device__ __forceinline__ HitR trace_wide(const DevScene& sc, const Staged& sg, const WideStack& ws, unsigned gmask, int j,
                                           const float4* po, const float4* pd, uint32_t ray_flags, bool any_hit) {
    Traversal T;
    if (!T.begin(sc, po, pd, ray_flags, any_hit)) return T.hit;
    const int gshift = __ffs(gmask) - 1;
    for (;;) {
        // ---- pop
        if (T.cur == 0) {
            bool empty = false;
            for (;;) {
                if (T.sp == 0) { empty = true; break; }
                const uint2 e = ws.at(--T.sp);
                T.cur = (int)e.x;
                if (T.cur > SENTINEL_KEEP) {
                    if (!(__uint_as_float(e.y) <= T.tcull)) continue;
                    if (T.bits & (TB_STALE | TB_STALE_T)) {
                        const float4 o = *po;
                        if (T.bits & TB_STALE) { const float4 d = *pd; T.set_ray(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z)); }
                        else T.set_origin(v3(o.x, o.y, o.z));
                        T.bits &= ~(TB_STALE | TB_STALE_T);
                    }
                    break;
                }
                T.ent = -1;
                if (T.cur == SENTINEL_RESTORE) T.bits |= TB_STALE;
                else if (T.cur == SENTINEL_RESTORE_T) T.bits |= TB_STALE_T;
            }
            if (empty) break;
            __syncwarp(gmask);   // every lane has read the popped entries before any lane pushes over them
        }
        if (T.cur > 0) {
            // ---- inner node: lane j tests child j
            const float* N = reinterpret_cast<const float*>(node_ptr<WHERE>(sc, sg, T.cur - 1));
            const int c = reinterpret_cast<const int*>(N)[48 + j];
            const float lx = N[j], hx = N[8 + j], ly = N[16 + j], hy = N[24 + j], lz = N[32 + j], hz = N[40 + j];
            const bool sx = (T.bits & TB_SX) != 0, sy = (T.bits & TB_SY) != 0, sz = (T.bits & TB_SZ) != 0;
            const float tn = fmaxf(fmaxf(fma_(T.idir.x, sx ? hx : lx, T.ilo.x), fma_(T.idir.y, sy ? hy : ly, T.ilo.y)), fmaxf(fma_(T.idir.z, sz ? hz : lz, T.ilo.z), T.tmin));
            const float tf = fminf(fminf(fma_(T.idir.x, sx ? lx : hx, T.ihi.x), fma_(T.idir.y, sy ? ly : hy, T.ihi.y)), fminf(fma_(T.idir.z, sz ? lz : hz, T.ihi.z), T.tcull));
            const bool hit = c != 0 && tn <= tf;
            const unsigned m = (__ballot_sync(gmask, hit) >> gshift) & 0xFFu;
            if (m == 0) { T.cur = 0; continue; }
            // nearest hit child: min over (tn, lane)
            float bt = hit ? tn : __int_as_float(0x7f800000); int bj = j;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const float ot = __shfl_xor_sync(gmask, bt, d); const int oj = __shfl_xor_sync(gmask, bj, d);
                if (ot < bt || (ot == bt && oj < bj)) { bt = ot; bj = oj; }
            }
            const unsigned others = m & ~(1u << bj);
            if (hit && j != bj) { const int pos = T.sp + __popc(others & ((1u << j) - 1u)); if (pos < WIDE_STACK) ws.at(pos) = make_uint2((uint32_t)c, __float_as_uint(tn)); }
            T.sp = min(T.sp + __popc(others), WIDE_STACK);
            T.cur = __shfl_sync(gmask, c, gshift + bj);
            __syncwarp(gmask);
        }
        if (T.cur < 0 && T.ent < 0) {
            // ---- entity: every lane runs the scalar step on its copy; one lane owns the stack write
            const int sp0 = T.sp;
            uint2 lst[1];
            Stack one; one.s = lst; one.stride = 0; one.l = lst;   // captures the single sentinel push of entity_step
            int spx = 0;
            {
                // entity_step pushes at most one entry (a sentinel) at index T.sp: redirect it into `lst`
                const int keep = T.sp; T.sp = 0;
                T.template entity_step<WHERE>(sc, sg, one);
                spx = T.sp; T.sp = keep;
            }
            if (spx > 0) { if (j == 0 && sp0 < WIDE_STACK) ws.at(sp0) = lst[0]; T.sp = min(sp0 + 1, WIDE_STACK); __syncwarp(gmask); }
        }
        if (T.cur < 0 && T.ent >= 0) {
            // ---- leaf: lane j tests triangle j
            const int r = -T.cur - 1;
            const int first = r >> 2, cnt = (r & 3) + 1;
            T.cur = 0;
            float t = __int_as_float(0x7f800000), u = 0, v = 0; int prim = -1;
            if (j < cnt) {
                const float4* P = tri_ptr<WHERE>(sc, sg, first + j);
                float tt, uu, vv;
                if (intersect_tri(T.org, T.dir, T.tmin, T.tmax, P[0], P[1], P[2], tt, uu, vv)) {
                    const int pp = __ldg(sc.tri_prim + first + j);
                    if (better(tt, T.ent, pp, T.hit)) { t = tt; u = uu; v = vv; prim = pp; }
                }
            }
            // best candidate of the group: smaller t, then larger primitive id (the order of `better`)
            float bt = t; int bp = prim, bj = j;
#pragma unroll
            for (int d = 1; d < 4; d <<= 1) {
                const float ot = __shfl_xor_sync(gmask, bt, d); const int op = __shfl_xor_sync(gmask, bp, d); const int oj = __shfl_xor_sync(gmask, bj, d);
                if (op >= 0 && (bp < 0 || ot < bt || (ot == bt && op > bp))) { bt = ot; bp = op; bj = oj; }
            }
            // lanes 0-3 agree; lanes 4-7 take the result from lane 0 of the group
            bj = __shfl_sync(gmask, bj, gshift); bp = __shfl_sync(gmask, bp, gshift);
            if (bp >= 0) {
                const float wt = __shfl_sync(gmask, t, gshift + bj), wu = __shfl_sync(gmask, u, gshift + bj), wv = __shfl_sync(gmask, v, gshift + bj);
                T.accept(wt, wu, wv, bp, T.ent);
                if (T.bits & TB_DONE) break;
            }
        }
        if (T.bits & TB_DONE) break;
    }
    return T.hit;
}

}  // namespace igb
