// bvh_build.hpp — host BVH builder for the device traversal (rt_device.cuh: BvhNode).
//
// Replaces bvh/bbox_tree/constructor.rs:9-212.  The reference builder scores six candidate
// partitions by summed child *volume*, rescans all N leaves per candidate per node (O(N^2))
// and degenerates beyond a few thousand objects (depth 644 at N=4098).
// Closest-hit results do not depend on the tree, so this is a different algorithm: top-down
// surface-area heuristic (exact sweep for small ranges, 32 bins otherwise), one primitive
// per leaf, children's boxes stored in the parent, nodes emitted breadth-first so the top
// of the tree is a prefix of the array.  Large ranges build their two halves as OpenMP tasks
// (1e6 spheres: ~1 s single-threaded).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace b200rt {

struct HostBox { float lo[3], hi[3]; };
struct HostNode { float q[16]; };   // same bytes as BvhNode (4 x float4)

struct BuildPrim { HostBox box; float centroid[3]; int code; };   // code = (type << 28) | id

struct BvhBuildResult {
    std::vector<HostNode> nodes;
    uint32_t depth = 0;             // max number of inner nodes on a root-to-leaf path
};

constexpr uint32_t BVH_MAX_DEPTH = 60;     // levels the device traversal stack holds (b200rt.cu checks the same bound)

namespace detail {
inline void box_init(HostBox& b) { for (int a = 0; a < 3; ++a) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; } }
inline void box_grow(HostBox& b, const HostBox& o) { for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); } }
inline float box_area(const HostBox& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.0f;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}
struct TmpNode { HostBox box[2]; int child[2]; };   // child >= 0: TmpNode index, < 0: ~code

struct Builder {
    std::vector<BuildPrim>& prims;
    std::vector<TmpNode> tmp;        // n - 1 inner nodes; the subtree of a range of k prims owns k - 1 consecutive slots
    uint32_t max_depth = 0;
    explicit Builder(std::vector<BuildPrim>& p) : prims(p), tmp(p.size() > 1 ? p.size() - 1 : 0) {}

    // Builds the subtree of range [b, e) into tmp[base ...]; returns its child reference.
    // Slots are assigned from the range sizes alone (self = base, left subtree next, then the
    // right one), so sibling subtrees are independent tasks and the result does not depend
    // on the thread count.
    int build(size_t b, size_t e, uint32_t depth, size_t base, HostBox* out_box, uint32_t* out_depth) {
        HostBox bb; box_init(bb);
        for (size_t i = b; i < e; ++i) box_grow(bb, prims[i].box);
        *out_box = bb;
        if (e - b == 1) { *out_depth = depth; return ~prims[b].code; }
        size_t mid = split(b, e, depth);
        HostBox lb, rb;
        uint32_t ld = 0, rd = 0;
        int l, r;
        const size_t n_left = mid - b;
        if (e - b >= 8192) {
#pragma omp task shared(l, lb, ld)
            l = build(b, mid, depth + 1, base + 1, &lb, &ld);
#pragma omp task shared(r, rb, rd)
            r = build(mid, e, depth + 1, base + n_left, &rb, &rd);
#pragma omp taskwait
        } else {
            l = build(b, mid, depth + 1, base + 1, &lb, &ld);
            r = build(mid, e, depth + 1, base + n_left, &rb, &rd);
        }
        tmp[base].box[0] = lb; tmp[base].box[1] = rb;
        tmp[base].child[0] = l; tmp[base].child[1] = r;
        *out_depth = std::max(std::max(ld, rd), depth + 1);
        return (int)base;
    }

    size_t split(size_t b, size_t e, uint32_t depth) {
        size_t n = e - b;
        HostBox cb; box_init(cb);
        for (size_t i = b; i < e; ++i) for (int a = 0; a < 3; ++a) { cb.lo[a] = std::min(cb.lo[a], prims[i].centroid[a]); cb.hi[a] = std::max(cb.hi[a], prims[i].centroid[a]); }
        float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
        int wide = ext[0] >= ext[1] ? (ext[0] >= ext[2] ? 0 : 2) : (ext[1] >= ext[2] ? 1 : 2);
        auto median_split = [&](int axis) {
            size_t mid = b + n / 2;
            std::nth_element(prims.begin() + b, prims.begin() + mid, prims.begin() + e,
                             [axis](const BuildPrim& x, const BuildPrim& y) { return x.centroid[axis] < y.centroid[axis] || (x.centroid[axis] == y.centroid[axis] && x.code < y.code); });
            return mid;
        };
        // Depth guard: the device traversal stack holds 60 levels.  A median split halves the range, so a subtree of n
        // primitives built by median splits alone is ceil(log2 n) deep: switch to them while that still fits (a SAH
        // split of a strongly clustered scene can peel one primitive per level).
        uint32_t log2n = 0; while ((size_t(1) << log2n) < n) ++log2n;
        if (!(ext[wide] > 0.0f) || depth + log2n >= BVH_MAX_DEPTH - 2) return median_split(wide);   // coincident centroids / depth guard

        float best_cost = INFINITY; int best_axis = -1; size_t best_mid = 0;
        if (n <= 64) {
            // exact sweep on each axis
            std::vector<float> right_area(n);
            for (int axis = 0; axis < 3; ++axis) {
                if (!(ext[axis] > 0.0f)) continue;
                std::sort(prims.begin() + b, prims.begin() + e, [axis](const BuildPrim& x, const BuildPrim& y) { return x.centroid[axis] < y.centroid[axis] || (x.centroid[axis] == y.centroid[axis] && x.code < y.code); });
                HostBox acc; box_init(acc);
                for (size_t i = n; i-- > 1;) { box_grow(acc, prims[b + i].box); right_area[i] = box_area(acc); }
                box_init(acc);
                for (size_t i = 1; i < n; ++i) {
                    box_grow(acc, prims[b + i - 1].box);
                    float cost = box_area(acc) * (float)i + right_area[i] * (float)(n - i);
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_mid = b + i; }
                }
            }
            if (best_axis < 0) return median_split(wide);
            std::sort(prims.begin() + b, prims.begin() + e, [best_axis](const BuildPrim& x, const BuildPrim& y) { int a = best_axis; return x.centroid[a] < y.centroid[a] || (x.centroid[a] == y.centroid[a] && x.code < y.code); });
            return best_mid;
        }
        const int NB = 32;
        int best_k = -1;
        auto bin_of = [&](const BuildPrim& x, int axis) {
            int k = (int)((x.centroid[axis] - cb.lo[axis]) * ((float)NB / ext[axis]));
            return std::min(std::max(k, 0), NB - 1);
        };
        for (int axis = 0; axis < 3; ++axis) {
            if (!(ext[axis] > 0.0f)) continue;
            HostBox bins[NB]; size_t cnt[NB];
            for (int k = 0; k < NB; ++k) { box_init(bins[k]); cnt[k] = 0; }
            for (size_t i = b; i < e; ++i) { int k = bin_of(prims[i], axis); box_grow(bins[k], prims[i].box); cnt[k]++; }
            float ra[NB]; size_t rc[NB];
            HostBox acc; box_init(acc); size_t c = 0;
            for (int k = NB - 1; k >= 1; --k) { box_grow(acc, bins[k]); c += cnt[k]; ra[k] = box_area(acc); rc[k] = c; }
            box_init(acc); c = 0;
            for (int k = 1; k < NB; ++k) {   // split between bins k-1 and k
                box_grow(acc, bins[k - 1]); c += cnt[k - 1];
                if (c == 0 || rc[k] == 0) continue;
                float cost = box_area(acc) * (float)c + ra[k] * (float)rc[k];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_k = k; }
            }
        }
        if (best_axis < 0) return median_split(wide);
        int axis = best_axis;
        auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](const BuildPrim& x) { return bin_of(x, axis) < best_k; });
        size_t mid = (size_t)(it - prims.begin());
        if (mid == b || mid == e) return median_split(wide);
        return mid;
    }
};
}  // namespace detail

// `prims`: one entry per primitive that can be hit (callers drop primitives whose box is
// inverted — e.g. negative-radius spheres, whose reference box never passes Aabb::hit2).
inline BvhBuildResult build_bvh(std::vector<BuildPrim> prims) {
    using namespace detail;
    BvhBuildResult res;
    const int EMPTY = (int)0x80000000;
    HostBox empty; box_init(empty);
    std::vector<TmpNode> tmp;
    uint32_t depth = 1;
    if (prims.empty()) {
        TmpNode t; t.box[0] = empty; t.box[1] = empty; t.child[0] = EMPTY; t.child[1] = EMPTY; tmp.push_back(t);
    } else if (prims.size() == 1) {
        TmpNode t; t.box[0] = prims[0].box; t.box[1] = empty; t.child[0] = ~prims[0].code; t.child[1] = EMPTY; tmp.push_back(t);
    } else {
        Builder bld(prims);
        HostBox rb;
        uint32_t d = 0;
#pragma omp parallel if (prims.size() >= 8192)
#pragma omp single nowait
        bld.build(0, prims.size(), 0, 0, &rb, &d);
        tmp.swap(bld.tmp);
        depth = d;
    }
    // breadth-first renumbering (root = 0 already, since build() allocates self before children)
    std::vector<int> order; order.reserve(tmp.size());
    std::vector<int> new_index(tmp.size(), -1);
    order.push_back(0);
    for (size_t head = 0; head < order.size(); ++head) {
        int t = order[head];
        new_index[t] = (int)head;
        for (int k = 0; k < 2; ++k) if (tmp[t].child[k] >= 0) order.push_back(tmp[t].child[k]);
    }
    res.nodes.resize(tmp.size());
    for (size_t i = 0; i < order.size(); ++i) {
        const TmpNode& t = tmp[order[i]];
        HostNode& n = res.nodes[i];
        const HostBox& a = t.box[0]; const HostBox& c = t.box[1];
        float q[16] = {a.lo[0], a.lo[1], a.lo[2], a.hi[0], a.hi[1], a.hi[2], c.lo[0], c.lo[1], c.lo[2], c.hi[0], c.hi[1], c.hi[2], 0, 0, 0, 0};
        int c0 = t.child[0] >= 0 ? new_index[t.child[0]] : t.child[0];
        int c1 = t.child[1] >= 0 ? new_index[t.child[1]] : t.child[1];
        std::memcpy(&q[12], &c0, 4); std::memcpy(&q[13], &c1, 4);
        std::memcpy(n.q, q, sizeof(q));
    }
    res.depth = depth;
    return res;
}

}  // namespace b200rt
