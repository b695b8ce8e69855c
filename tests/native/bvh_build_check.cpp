// Host-side check of the binned-SAH builder (shirley_raytracing_rs_b200/csrc/bvh_build.hpp), compiled and run by
// tests/test_bvh_build.py: tree depth stays inside the device traversal stack on adversarial inputs, every primitive
// is a leaf exactly once, and every child box contains the boxes below it.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../../shirley_raytracing_rs_b200/csrc/bvh_build.hpp"

using namespace b200rt;

static BuildPrim sphere(float x, float y, float z, float r, int id) {
    BuildPrim p;
    p.box.lo[0] = x - r; p.box.lo[1] = y - r; p.box.lo[2] = z - r;
    p.box.hi[0] = x + r; p.box.hi[1] = y + r; p.box.hi[2] = z + r;
    p.centroid[0] = x; p.centroid[1] = y; p.centroid[2] = z;
    p.code = id;
    return p;
}

// returns the number of leaves under `ref`, checks containment; *depth = inner nodes on the longest path
static size_t walk(const std::vector<HostNode>& nodes, int ref, const HostBox& bound, std::vector<int>& seen, uint32_t level, uint32_t* depth, bool* ok) {
    if (ref < 0) {
        if (ref == (int)0x80000000) return 0;
        int id = ~ref;
        if (id < 0 || id >= (int)seen.size()) { *ok = false; return 0; }
        seen[id]++;
        if (level > *depth) *depth = level;
        return 1;
    }
    const float* q = nodes[ref].q;
    size_t n = 0;
    for (int ch = 0; ch < 2; ++ch) {
        HostBox b; for (int k = 0; k < 3; ++k) { b.lo[k] = q[6 * ch + k]; b.hi[k] = q[6 * ch + 3 + k]; }
        int c; std::memcpy(&c, &q[12 + ch], 4);
        if (c != (int)0x80000000) for (int k = 0; k < 3; ++k) if (b.lo[k] < bound.lo[k] || b.hi[k] > bound.hi[k]) *ok = false;
        n += walk(nodes, c, b, seen, level + 1, depth, ok);
    }
    return n;
}

static int check(const char* name, std::vector<BuildPrim> prims) {
    const size_t n = prims.size();
    std::vector<BuildPrim> copy = prims;
    BvhBuildResult r = build_bvh(std::move(prims));
    HostBox all; detail::box_init(all);
    for (auto& p : copy) detail::box_grow(all, p.box);
    std::vector<int> seen(n, 0);
    uint32_t depth = 0; bool ok = true;
    size_t leaves = walk(r.nodes, 0, all, seen, 0, &depth, &ok);
    for (size_t i = 0; i < n; ++i) if (seen[i] != 1) ok = false;
    // leaf boxes equal the primitives' boxes
    printf("%s: n %zu nodes %zu depth %u (reported %u) leaves %zu %s\n", name, n, r.nodes.size(), depth, r.depth, leaves, ok ? "ok" : "BAD");
    if (!ok || leaves != n || depth != r.depth || r.depth > BVH_MAX_DEPTH || (n >= 2 && r.nodes.size() != n - 1)) return 1;
    return 0;
}

int main() {
    int bad = 0;
    {   // geometric cluster: sizes and positions double, a surface-area split peels one primitive per level
        std::vector<BuildPrim> p;
        float x = 1.0f;
        for (int k = 0; k < 120; ++k) { p.push_back(sphere(x, 0.f, 0.f, 0.2f * x, k)); x *= 2.0f; }
        bad += check("geometric x2, 120", p);
    }
    {   // the same, nested in three dimensions, 3000 primitives (ratio 1.01 .. strong)
        std::vector<BuildPrim> p;
        std::mt19937 rng(5);
        std::uniform_real_distribution<float> u(-1.f, 1.f);
        float s = 1e-6f;
        for (int k = 0; k < 3000; ++k) { p.push_back(sphere(s * u(rng), s * u(rng), s * u(rng), 0.3f * s, k)); s *= 1.02f; }
        bad += check("nested shells, 3000", p);
    }
    {   // coincident centroids (all splits degenerate)
        std::vector<BuildPrim> p;
        for (int k = 0; k < 1000; ++k) p.push_back(sphere(1.f, 2.f, 3.f, 0.5f + 0.001f * k, k));
        bad += check("coincident, 1000", p);
    }
    {   // a plain random scene for reference
        std::vector<BuildPrim> p;
        std::mt19937 rng(7);
        std::uniform_real_distribution<float> u(-100.f, 100.f);
        for (int k = 0; k < 20000; ++k) p.push_back(sphere(u(rng), 0.2f, u(rng), 0.2f, k));
        bad += check("random, 20000", p);
    }
    {   // one long line with exponentially growing gaps AND many primitives: depth guard + large n
        std::vector<BuildPrim> p;
        float x = 1.0f;
        for (int k = 0; k < 100000; ++k) { p.push_back(sphere(x, 0.f, 0.f, 1e-4f * x, k)); x *= 1.0005f; }
        bad += check("geometric x1.0005, 100000", p);
    }
    return bad ? 1 : 0;
}
