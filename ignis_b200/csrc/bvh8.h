// BVH8 construction on the host: binned-SAH binary build, then collapse to arity 8.
//
// Replaces what the reference gets from the external madmann91/bvh builder plus its N-ary collapse
// (src/runtime/bvh/NArityBvh.h:94-143, BvhNAdapter.h:38-93, TriBVHAdapter.h:196-223, SceneBVHAdapter.h:110-128).
// The node keeps the reference's `Node8` shape (bounds[6][8] = lo_x,hi_x,lo_y,hi_y,lo_z,hi_z per child lane,
// then child[8], src/artic/traversal/bvh.art:86-90) because that shape is exactly eight float4 pairs, i.e. two
// 128-byte lines per node; the child encoding is this device's own:
//   child > 0 : inner node, index child-1 (relative to the tree's first node)
//   child < 0 : leaf, r = -child-1, first primitive slot = r >> 2, primitive count = (r & 3) + 1
//   child = 0 : empty lane (children are packed to the front); its box is inverted with finite bounds (+-FLT_MAX)
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace igb {

struct Box3 {
    float lo[3], hi[3];
    static Box3 empty() {
        const float inf = std::numeric_limits<float>::infinity();
        return Box3{{inf, inf, inf}, {-inf, -inf, -inf}};
    }
    void extend(const Box3& b) {
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    }
    void extend(const float p[3]) {
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.f;
        return dx * dy + dx * dz + dy * dz;
    }
};

struct alignas(16) Node8 {
    float   bounds[6][8];
    int32_t child[8];
    int32_t pad[8];
};
static_assert(sizeof(Node8) == 256, "Node8 must be two 128-byte lines");

struct Bvh8 {
    std::vector<Node8>   nodes;  // nodes[0] is the root
    std::vector<int32_t> order;  // primitive slot -> input primitive index (leaves reference contiguous slots)
    int                  max_depth = 0;
};

namespace detail {

struct Bvh2Node { Box3 box; int left, right, first, count; };

struct Builder2 {
    const std::vector<Box3>& boxes;
    std::vector<float>       cen;    // 3 per prim
    std::vector<int>         order;
    std::vector<Bvh2Node>    nodes;
    int                      max_leaf;

    Builder2(const std::vector<Box3>& b, int leaf) : boxes(b), max_leaf(leaf) {
        const int n = (int)b.size();
        cen.resize(3 * (size_t)n);
        order.resize(n);
        for (int i = 0; i < n; ++i) {
            order[i] = i;
            for (int k = 0; k < 3; ++k) cen[3 * i + k] = 0.5f * (b[i].lo[k] + b[i].hi[k]);
        }
        nodes.reserve(2 * (size_t)n + 1);
        if (n) build(0, n);
    }

    int build(int b, int e) {
        const int id = (int)nodes.size();
        nodes.push_back(Bvh2Node{});
        Box3 bb = Box3::empty(), cb = Box3::empty();
        for (int i = b; i < e; ++i) { bb.extend(boxes[order[i]]); cb.extend(&cen[3 * order[i]]); }
        nodes[id].box = bb;
        const int n = e - b;
        auto make_leaf = [&]() { nodes[id].first = b; nodes[id].count = n; nodes[id].left = nodes[id].right = -1; return id; };
        // SIMT traversal pays a fixed price per visit, so leaves are filled up to max_leaf primitives (no SAH leaf test)
        if (n <= max_leaf) return make_leaf();
        // binned SAH over the three axes
        constexpr int NB = 16;
        float best_cost = std::numeric_limits<float>::infinity();
        int best_axis = -1, best_bin = -1;
        for (int axis = 0; axis < 3; ++axis) {
            const float lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0)) continue;
            Box3 bin_box[NB]; int bin_cnt[NB];
            for (int k = 0; k < NB; ++k) { bin_box[k] = Box3::empty(); bin_cnt[k] = 0; }
            const float scale = NB / ext;
            for (int i = b; i < e; ++i) {
                int k = (int)((cen[3 * order[i] + axis] - lo) * scale);
                k = std::min(std::max(k, 0), NB - 1);
                bin_box[k].extend(boxes[order[i]]); bin_cnt[k]++;
            }
            float right_area[NB]; int right_cnt[NB];
            Box3 acc = Box3::empty(); int cnt = 0;
            for (int k = NB - 1; k > 0; --k) { acc.extend(bin_box[k]); cnt += bin_cnt[k]; right_area[k] = acc.half_area(); right_cnt[k] = cnt; }
            acc = Box3::empty(); cnt = 0;
            for (int k = 0; k < NB - 1; ++k) {
                acc.extend(bin_box[k]); cnt += bin_cnt[k];
                if (cnt == 0 || right_cnt[k + 1] == 0) continue;
                const float cost = acc.half_area() * (float)cnt + right_area[k + 1] * (float)right_cnt[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = k; }
            }
        }
        int mid;
        if (best_axis < 0) {
            mid = (b + e) / 2;  // identical centroids: split in input order
        } else {
            const float lo = cb.lo[best_axis], scale = NB / (cb.hi[best_axis] - cb.lo[best_axis]);
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](int p) {
                int k = (int)((cen[3 * p + best_axis] - lo) * scale);
                k = std::min(std::max(k, 0), NB - 1);
                return k <= best_bin;
            });
            mid = (int)(it - order.begin());
            if (mid == b || mid == e) mid = (b + e) / 2;
        }
        nodes[id].count = 0; nodes[id].first = 0;
        const int l = build(b, mid);
        const int r = build(mid, e);
        nodes[id].left = l; nodes[id].right = r;
        return id;
    }
};

}  // namespace detail

// Builds a BVH8 over `boxes`. max_leaf in [1,4].
inline Bvh8 build_bvh8(const std::vector<Box3>& boxes, int max_leaf) {
    Bvh8 out;
    if (boxes.empty()) return out;
    max_leaf = std::min(std::max(max_leaf, 1), 4);
    detail::Builder2 b2(boxes, max_leaf);
    out.order = b2.order;
    const auto& n2 = b2.nodes;
    // empty child lanes hold an inverted box with FINITE bounds: the slab test then misses without a `child != 0`
    // check and without 0 * inf / inf - inf NaNs (traverse.cuh)
    const float inf = std::numeric_limits<float>::max();

    struct Work { int n2; int n8; int depth; };
    std::vector<Work> stack;
    auto new_node = [&]() {
        Node8 n;
        for (int k = 0; k < 6; ++k) for (int c = 0; c < 8; ++c) n.bounds[k][c] = (k & 1) ? -inf : inf;
        for (int c = 0; c < 8; ++c) { n.child[c] = 0; n.pad[c] = 0; }
        out.nodes.push_back(n);
        return (int)out.nodes.size() - 1;
    };
    const int root = new_node();
    stack.push_back(Work{0, root, 1});
    while (!stack.empty()) {
        const Work w = stack.back();
        stack.pop_back();
        out.max_depth = std::max(out.max_depth, w.depth);
        int kids[8]; int nk = 0;
        if (n2[w.n2].count > 0) { kids[nk++] = w.n2; }
        else {
            kids[nk++] = n2[w.n2].left; kids[nk++] = n2[w.n2].right;
            while (nk < 8) {  // open the inner child with the largest surface area
                int best = -1; float best_area = -1.f;
                for (int i = 0; i < nk; ++i)
                    if (n2[kids[i]].count == 0) { const float a = n2[kids[i]].box.half_area(); if (a > best_area) { best_area = a; best = i; } }
                if (best < 0) break;
                const int k = kids[best];
                kids[best] = n2[k].left; kids[nk++] = n2[k].right;
            }
        }
        for (int i = 0; i < nk; ++i) {
            const detail::Bvh2Node& c = n2[kids[i]];
            Node8& n = out.nodes[w.n8];
            n.bounds[0][i] = c.box.lo[0]; n.bounds[1][i] = c.box.hi[0];
            n.bounds[2][i] = c.box.lo[1]; n.bounds[3][i] = c.box.hi[1];
            n.bounds[4][i] = c.box.lo[2]; n.bounds[5][i] = c.box.hi[2];
            if (c.count > 0) {
                // leaves of the binary build hold <= max_leaf primitives, except degenerate runs: chunk those
                int first = c.first, left = c.count;
                // (a run longer than 4 can only come from identical centroids with max_leaf honoured, so never here)
                n.child[i] = -(((first << 2) | (std::min(left, 4) - 1)) + 1);
            } else {
                const int id = new_node();
                out.nodes[w.n8].child[i] = id + 1;
                stack.push_back(Work{kids[i], id, w.depth + 1});
            }
        }
    }
    return out;
}

}  // namespace igb
