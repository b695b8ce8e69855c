// oracle.hpp — CPU restatement of the reference's per-pixel path-tracing loop.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (shirley_raytracing_rs_b200/, the C ABI
// in include/b200rt.h) includes, links or calls this.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use it, as the checker or as the
// timed CPU baseline.
//
// What it restates: scottschroeder/shirley-raytracing-rs, file:line cited at each function
// (paths relative to /root/reference/src/raytracer unless stated).  The reference is pure
// Rust and cannot be built in this image (no rustc/cargo, 143 un-vendored crates), so this
// is a "port" oracle.  PARITY PINNING: the reference's own unit tests for this path
// (bvh/aabb.rs:94-166, bvh/bbox_tree.rs:103-227, core/fp.rs:35-112 — SURVEY.md §4 KA1-KA9)
// are replayed verbatim against this code in tests/test_oracle_kat.py; nothing else in the
// reference pins camera rays, scatter, textures or rendered images (its RNG is unseeded),
// so those layers are pinned only by being line-by-line restatements.
// Third-party arithmetic restated from published behaviour (sources absent from
// /root/reference): nalgebra 0.31.1 Vector3 ops (dot = (x*x'+y*y')+z*z', normalize divides
// by the norm), rand 0.8.5 gen::<f64>() (uniform [0,1)) — replaced by the counter-based
// generator documented in include/b200rt.h so GPU and oracle can share injected randoms.
//
// template<class Real>: Real = double is the reference-faithful mode (and the CPU
// baseline); Real = float runs the same algorithm in f32.  Compile with -ffp-contract=off:
// rustc never fuses a*b+c.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/b200rt.h"

namespace oracle {

// ------------------------------------------------------------------------------------------
// RNG (documented in include/b200rt.h: b200rt_rng_uniforms).  Replaces rand::ThreadRng
// (src/main.rs:98,119); the reference's stream is OS-seeded and not reproducible.
// ------------------------------------------------------------------------------------------
inline uint32_t pcg_out(uint32_t s) {
    uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
inline uint32_t hash32(uint32_t x) { return pcg_out(x * 747796405u + 2891336453u); }

struct Rng {
    uint32_t state = 0, inc = 1;
    uint32_t draws = 0;
    Rng() = default;
    Rng(uint64_t seed, uint32_t a, uint32_t b) {
        uint32_t k0 = hash32((uint32_t)seed);
        uint32_t k1 = hash32((uint32_t)(seed >> 32) ^ k0);
        state = hash32(b + hash32(a ^ k1));
        inc = (hash32(a + hash32(b ^ k0)) << 1) | 1u;
    }
    uint32_t next_u32() {
        state = state * 747796405u + inc;
        ++draws;
        return pcg_out(state);
    }
    // rng.gen::<f64>() — core/math.rs:24.  24-bit resolution so the same value is exact in
    // f32 and f64.
    template <class Real> Real gen() { return (Real)(next_u32() >> 8) * (Real)(1.0 / 16777216.0); }
};

// A source of uniforms: either the generator above or a caller-provided list (injected
// randoms for the scatter / camera parity tests).
struct UniformSource {
    Rng rng;
    const double* injected = nullptr;
    size_t n_injected = 0, pos = 0;
    template <class Real> Real gen() {
        if (injected) {
            double v = pos < n_injected ? injected[pos] : 0.5;
            ++pos;
            return (Real)v;
        }
        return rng.gen<Real>();
    }
    uint32_t draws() const { return injected ? (uint32_t)pos : rng.draws; }
};

// ------------------------------------------------------------------------------------------
// core/fp.rs:3-28 — NaN-aware min / max
// ------------------------------------------------------------------------------------------
template <class Real> inline Real non_nan(Real a, Real b) { return std::isnan(a) ? b : a; }
template <class Real> inline Real fmin_(Real a, Real b) {
    if (std::isnan(a) || std::isnan(b)) return non_nan(a, b);
    return a < b ? a : b;   // Some(Less) => a, Some(_) => b
}
template <class Real> inline Real fmax_(Real a, Real b) {
    if (std::isnan(a) || std::isnan(b)) return non_nan(a, b);
    return a > b ? a : b;   // Some(Greater) => a, Some(_) => b
}

// ------------------------------------------------------------------------------------------
// core/vec3.rs:84-229 over nalgebra::Vector3
// ------------------------------------------------------------------------------------------
template <class Real> struct Vec3 {
    Real x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(Real x_, Real y_, Real z_) : x(x_), y(y_), z(z_) {}
    Real operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Real& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
    Vec3 operator+(const Vec3& r) const { return {x + r.x, y + r.y, z + r.z}; }
    Vec3 operator-(const Vec3& r) const { return {x - r.x, y - r.y, z - r.z}; }
    Vec3 operator*(const Vec3& r) const { return {x * r.x, y * r.y, z * r.z}; }   // vec3.rs:198-208
    Vec3 scale(Real s) const { return {x * s, y * s, z * s}; }                   // vec3.rs:123-127
    Real dot(const Vec3& r) const { return (x * r.x + y * r.y) + z * r.z; }      // nalgebra dot
    Real length_squared() const { return dot(*this); }                            // vec3.rs:118
    Real length() const { return std::sqrt(length_squared()); }                   // vec3.rs:113
    Vec3 cross(const Vec3& r) const {                                             // vec3.rs:162
        return {y * r.z - z * r.y, z * r.x - x * r.z, x * r.y - y * r.x};
    }
    Vec3 unit() const {                                                           // vec3.rs:170
        Real n = length();
        return {x / n, y / n, z / n};
    }
    bool near_zero() const {                                                      // vec3.rs:130
        const Real e = (Real)1e-8;
        return std::fabs(x) < e && std::fabs(y) < e && std::fabs(z) < e;
    }
    Vec3 reflect(const Vec3& n) const { return *this - n.scale((Real)2 * dot(n)); }   // vec3.rs:135
    Vec3 refract(const Vec3& n, Real etai_over_etat) const {                           // vec3.rs:139-145
        Real cos_theta = fmin_<Real>(scale((Real)-1).dot(n), (Real)1);
        Vec3 r_out_perp = (n.scale(cos_theta) + *this).scale(etai_over_etat);
        Real r_out_parallel_mag = std::sqrt(std::fabs((Real)1 - r_out_perp.length_squared())) * (Real)-1;
        Vec3 r_out_parallel = n.scale(r_out_parallel_mag);
        return r_out_perp + r_out_parallel;
    }
};

template <class Real> struct Ray {                                                // vec3.rs:240-256
    Vec3<Real> orig, direction;
    Vec3<Real> at(Real t) const { return orig + direction.scale(t); }
};

// ------------------------------------------------------------------------------------------
// bvh/aabb.rs
// ------------------------------------------------------------------------------------------
template <class Real> struct Aabb {
    Vec3<Real> min, max;
    // aabb.rs:62-79 (the variant BboxTree uses)
    bool hit2(const Ray<Real>& r, Real t_min, Real t_max) const {
        for (int a = 0; a < 3; ++a) {
            Real inv_d = (Real)1 / r.direction[a];
            Real t0 = (min[a] - r.orig[a]) * inv_d;
            Real t1 = (max[a] - r.orig[a]) * inv_d;
            if (inv_d < (Real)0) std::swap(t0, t1);
            t_min = t0 > t_min ? t0 : t_min;
            t_max = t1 < t_max ? t1 : t_max;
            if (t_max <= t_min) return false;
        }
        return true;
    }
    // aabb.rs:42-61 (unused variant, kept for the KATs)
    bool hit(const Ray<Real>& r, Real t_min, Real t_max) const {
        for (int a = 0; a < 3; ++a) {
            Real ta = (min[a] - r.orig[a]) / r.direction[a];
            Real tb = (max[a] - r.orig[a]) / r.direction[a];
            Real t0 = fmin_<Real>(ta, tb), t1 = fmax_<Real>(ta, tb);
            t_min = fmax_<Real>(t0, t_min);
            t_max = fmin_<Real>(t1, t_max);
            if (t_max <= t_min) return false;
        }
        return true;
    }
    Real area() const {   // aabb.rs:81-86 (a volume, despite the name)
        return (max.x - min.x) * (max.y - min.y) * (max.z - min.z);
    }
};
template <class Real> inline Aabb<Real> surrounding_box(const Aabb<Real>& l, const Aabb<Real>& r) {   // aabb.rs:18-33
    Aabb<Real> o;
    o.min = {fmin_<Real>(l.min.x, r.min.x), fmin_<Real>(l.min.y, r.min.y), fmin_<Real>(l.min.z, r.min.z)};
    o.max = {fmax_<Real>(l.max.x, r.max.x), fmax_<Real>(l.max.y, r.max.y), fmax_<Real>(l.max.z, r.max.z)};
    return o;
}

// ------------------------------------------------------------------------------------------
// geometry/
// ------------------------------------------------------------------------------------------
template <class Real> struct HitRecord {   // hittable.rs:6-38
    Vec3<Real> point, normal;
    Real t = 0, u = 0, v = 0;
    bool front_face = false;
    static HitRecord make(const Ray<Real>& incoming, Vec3<Real> point, Vec3<Real> normal, Real t, Real u, Real v) {
        HitRecord h;
        h.front_face = incoming.direction.dot(normal) < (Real)0;
        if (!h.front_face) normal = normal.scale((Real)-1);
        h.point = point; h.normal = normal; h.t = t; h.u = u; h.v = v;
        return h;
    }
};

template <class Real> struct Sphere {      // sphere.rs
    Vec3<Real> center; Real radius;
    static void get_uv(const Vec3<Real>& p, Real& u, Real& v) {   // sphere.rs:18-25
        const Real PI = (Real)3.14159265358979323846;
        Real theta = std::acos(-p.y);
        Real phi = std::atan2(-p.z, p.x) + PI;
        u = phi / ((Real)2 * PI);
        v = theta / PI;
    }
    bool hit(const Ray<Real>& ray, Real t_min, Real t_max, HitRecord<Real>& out) const {   // sphere.rs:29-52
        Vec3<Real> oc = ray.orig - center;
        Real a = ray.direction.length_squared();
        Real half_b = oc.dot(ray.direction);
        Real c = oc.length_squared() - radius * radius;
        Real discriminant = half_b * half_b - a * c;
        if (discriminant < (Real)0) return false;
        Real sqrt_d = std::sqrt(discriminant);
        Real root = (-half_b - sqrt_d) / a;
        if (root < t_min || t_max < root) {
            root = (-half_b + sqrt_d) / a;
            if (root < t_min || t_max < root) return false;
        }
        Vec3<Real> point = ray.at(root);
        Vec3<Real> normal = (point - center).scale((Real)1 / radius);
        Real u, v;
        get_uv(normal, u, v);
        out = HitRecord<Real>::make(ray, point, normal, root, u, v);
        return true;
    }
    Aabb<Real> bounding_box() const {      // sphere.rs:54-60
        Vec3<Real> r(radius, radius, radius);
        return {center - r, center + r};
    }
};

template <class Real> struct Rect {        // rect.rs:45-100; d1/d2 are the in-plane axes
    int d1 = 0, d2 = 1;
    Real d1_min = 0, d1_max = 0, d2_min = 0, d2_max = 0, offset = 0;
    bool hit(const Ray<Real>& ray, Real t_min, Real t_max, HitRecord<Real>& out) const {   // rect.rs:55-80
        int dn = 3 - d1 - d2;
        Real t = (offset - ray.orig[dn]) / ray.direction[dn];
        if (t < t_min || t > t_max) return false;
        Real d1v = ray.orig[d1] + t * ray.direction[d1];
        Real d2v = ray.orig[d2] + t * ray.direction[d2];
        if (d1v < d1_min || d1v > d1_max || d2v < d2_min || d2v > d2_max) return false;
        Real u = (d1v - d1_min) / (d1_max - d1_min);
        Real v = (d2v - d2_min) / (d2_max - d2_min);
        Vec3<Real> normal;
        normal.at(dn) = (Real)1;
        out = HitRecord<Real>::make(ray, ray.at(t), normal, t, u, v);
        return true;
    }
    Aabb<Real> bounding_box() const {      // rect.rs:82-99, BBOX_WIDTH rect.rs:9
        const Real W = (Real)0.0001;
        int dn = 3 - d1 - d2;
        Aabb<Real> b;
        b.min.at(d1) = d1_min; b.min.at(d2) = d2_min; b.min.at(dn) = offset - W;
        b.max.at(d1) = d1_max; b.max.at(d2) = d2_max; b.max.at(dn) = offset + W;
        return b;
    }
};
template <class Real> inline Rect<Real> make_rect(int d1, int d2, Real a, Real b, Real c, Real d, Real k) {
    Rect<Real> r; r.d1 = d1; r.d2 = d2; r.d1_min = a; r.d1_max = b; r.d2_min = c; r.d2_max = d; r.offset = k;
    return r;
}

template <class Real> struct RectBox {     // rect.rs:102-163
    Vec3<Real> min, max;
    Rect<Real> sides[6];                   // xy[0], xy[1], yz[0], yz[1], xz[0], xz[1]
    RectBox() = default;
    RectBox(Vec3<Real> p0, Vec3<Real> p1) : min(p0), max(p1) {   // rect.rs:112-129
        sides[0] = make_rect<Real>(0, 1, p0.x, p1.x, p0.y, p1.y, p1.z);
        sides[1] = make_rect<Real>(0, 1, p0.x, p1.x, p0.y, p1.y, p0.z);
        sides[2] = make_rect<Real>(1, 2, p0.y, p1.y, p0.z, p1.z, p1.x);
        sides[3] = make_rect<Real>(1, 2, p0.y, p1.y, p0.z, p1.z, p0.x);
        sides[4] = make_rect<Real>(0, 2, p0.x, p1.x, p0.z, p1.z, p1.y);
        sides[5] = make_rect<Real>(0, 2, p0.x, p1.x, p0.z, p1.z, p0.y);
    }
    bool hit(const Ray<Real>& ray, Real t_min, Real t_max, HitRecord<Real>& out) const {   // rect.rs:132-156
        bool any = false;
        for (int s = 0; s < 6; ++s) {
            Real t_closest = any ? out.t : t_max;      // check_closer, rect.rs:139
            HitRecord<Real> h;
            if (sides[s].hit(ray, t_min, t_closest, h)) { out = h; any = true; }
        }
        return any;
    }
    Aabb<Real> bounding_box() const { return {min, max}; }   // rect.rs:158-163
};

// geometry/object.rs:9-69 — closed enum dispatch
template <class Real> struct GeometricObject {
    uint32_t type = B200RT_PRIM_SPHERE;
    Sphere<Real> sphere{};
    Rect<Real> rect{};
    RectBox<Real> box{};
    bool hit(const Ray<Real>& ray, Real t_min, Real t_max, HitRecord<Real>& out) const {
        switch (type) {
            case B200RT_PRIM_SPHERE: return sphere.hit(ray, t_min, t_max, out);
            case B200RT_PRIM_BOX: return box.hit(ray, t_min, t_max, out);
            default: return rect.hit(ray, t_min, t_max, out);
        }
    }
    Aabb<Real> bounding_box() const {
        switch (type) {
            case B200RT_PRIM_SPHERE: return sphere.bounding_box();
            case B200RT_PRIM_BOX: return box.bounding_box();
            default: return rect.bounding_box();
        }
    }
};

// ------------------------------------------------------------------------------------------
// bvh/bbox_tree.rs + bvh/bbox_tree/constructor.rs
// ------------------------------------------------------------------------------------------
template <class Real> struct TreeNode {   // bbox_tree.rs:10-20: Aabb + enum NodePointer { Branch{lhs,rhs}, Leaf(idx) }
    Aabb<Real> bbox;
    bool leaf = true;             // enum tag (8 bytes with padding)
    size_t lhs = 0, rhs = 0;      // Branch { lhs, rhs };  Leaf(idx) keeps idx in `lhs`
    size_t leaf_idx() const { return lhs; }
};
static_assert(sizeof(TreeNode<double>) == 72, "bbox_tree.rs:103-107 size_of::<TreeNode>() == 72");

template <class Real> struct BboxTree {
    bool has_root = false;
    size_t root = 0;
    std::vector<TreeNode<Real>> tree;
    size_t max_stack = 0;   // statistics only

    // ---- constructor.rs ---------------------------------------------------------------
    struct Order { std::vector<std::pair<size_t, Real>> v; };
    static Order sorted_with_idx(const std::vector<TreeNode<Real>>& nodes, int axis) {   // constructor.rs:49-53
        Order o;
        for (size_t i = 0; i < nodes.size(); ++i) o.v.push_back({i, nodes[i].bbox.min[axis]});
        // sort_unstable_by(total_cmp): tie order is unspecified in the reference; a stable
        // sort (ties by index) is one valid outcome.
        std::stable_sort(o.v.begin(), o.v.end(), [](auto& a, auto& b) { return a.second < b.second; });
        return o;
    }
    typedef std::vector<char> BoxSet;   // membership bitmap, stands in for HashSet<usize>
    static size_t set_len(const BoxSet& s) { size_t n = 0; for (char c : s) n += c; return n; }

    static void split_median(const BoxSet& in, size_t in_len, const Order& order, BoxSet& l, BoxSet& r) {   // constructor.rs:77-93
        l.assign(in.size(), 0); r.assign(in.size(), 0);
        size_t ln = 0;
        for (auto& e : order.v) {
            if (!in[e.first]) continue;
            if (ln < in_len / 2) { l[e.first] = 1; ++ln; } else r[e.first] = 1;
        }
    }
    static void split_space(const BoxSet& in, const Order& order, BoxSet& l, BoxSet& r) {   // constructor.rs:95-124
        l.assign(in.size(), 0); r.assign(in.size(), 0);
        bool have_first = false; Real first = 0, last = 0;
        for (auto& e : order.v) if (in[e.first]) { if (!have_first) { first = e.second; have_first = true; } last = e.second; }
        Real mid = (last + first) / (Real)2;
        bool is_first = true;
        for (auto& e : order.v) {
            if (!in[e.first]) continue;
            if (is_first) { l[e.first] = 1; is_first = false; continue; }
            if (e.second < mid) l[e.first] = 1; else r[e.first] = 1;
        }
    }
    static Real total_area(const std::vector<TreeNode<Real>>& nodes, const BoxSet& l, const BoxSet& r) {   // constructor.rs:126-134
        auto vol = [&](const BoxSet& s) -> Real {
            bool any = false; Aabb<Real> b;
            for (size_t i = 0; i < s.size(); ++i) if (s[i]) { b = any ? surrounding_box(b, nodes[i].bbox) : nodes[i].bbox; any = true; }
            return any ? b.area() : (Real)0;
        };
        Real la = vol(l);
        Real ra = vol(r);
        return la + ra;
    }
    // f64::total_cmp for the split scores
    static bool total_less(Real a, Real b) {
        if (std::isnan(a) || std::isnan(b)) {
            // total order: -NaN < ... < +NaN; scores here are never NaN for finite scenes
            bool an = std::isnan(a), bn = std::isnan(b);
            if (an && bn) return std::signbit(a) && !std::signbit(b);
            if (an) return std::signbit(a);
            return !std::signbit(b);
        }
        if (a == b) return std::signbit(a) && !std::signbit(b);
        return a < b;
    }
    TreeNode<Real> partition_nodes(const std::vector<TreeNode<Real>>& nodes, const BoxSet& ws, const Order order[3]) {   // constructor.rs:182-212
        size_t n = set_len(ws);
        if (n == 1) { for (size_t i = 0; i < ws.size(); ++i) if (ws[i]) return nodes[i]; }
        // split_best, constructor.rs:151-180: x median, x space, y median, y space, z median, z space
        BoxSet bl, br, l, r;
        Real best = 0; bool have = false;
        for (int c = 0; c < 6; ++c) {
            if (c % 2 == 0) split_median(ws, n, order[c / 2], l, r); else split_space(ws, order[c / 2], l, r);
            Real score = total_area(nodes, l, r);
            if (!have || total_less(score, best)) { best = score; bl = l; br = r; have = true; }   // min_by keeps the FIRST minimum
        }
        TreeNode<Real> lhs = partition_nodes(nodes, bl, order);
        TreeNode<Real> rhs = partition_nodes(nodes, br, order);
        TreeNode<Real> out;
        out.bbox = surrounding_box(lhs.bbox, rhs.bbox);
        out.leaf = false;
        out.lhs = tree.size(); tree.push_back(lhs);
        out.rhs = tree.size(); tree.push_back(rhs);
        return out;
    }
    // constructor.rs:9-36.  reference_topology = false switches to a plain median split on
    // the widest axis: the reference builder is O(N^2) and degenerates (depth in the
    // hundreds) beyond a few thousand objects; closest-hit results do not depend on
    // topology.
    void construct(const std::vector<Aabb<Real>>& boxes, bool reference_topology) {
        tree.clear(); has_root = false;
        if (boxes.empty()) return;
        std::vector<TreeNode<Real>> leaves(boxes.size());
        for (size_t i = 0; i < boxes.size(); ++i) { leaves[i].bbox = boxes[i]; leaves[i].leaf = true; leaves[i].lhs = i; }
        TreeNode<Real> root_node;
        if (reference_topology) {
            Order order[3] = {sorted_with_idx(leaves, 0), sorted_with_idx(leaves, 1), sorted_with_idx(leaves, 2)};
            BoxSet all(leaves.size(), 1);
            root_node = partition_nodes(leaves, all, order);
        } else {
            std::vector<size_t> ids(leaves.size());
            for (size_t i = 0; i < ids.size(); ++i) ids[i] = i;
            root_node = median_build(leaves, ids.data(), ids.size());
        }
        root = tree.size(); has_root = true;
        tree.push_back(root_node);
    }
    TreeNode<Real> median_build(const std::vector<TreeNode<Real>>& leaves, size_t* ids, size_t n) {
        if (n == 1) return leaves[ids[0]];
        Aabb<Real> cb = leaves[ids[0]].bbox;
        for (size_t i = 1; i < n; ++i) cb = surrounding_box(cb, leaves[ids[i]].bbox);
        int axis = 0;
        Real ex = cb.max.x - cb.min.x, ey = cb.max.y - cb.min.y, ez = cb.max.z - cb.min.z;
        if (ey > ex && ey >= ez) axis = 1; else if (ez > ex && ez > ey) axis = 2;
        std::nth_element(ids, ids + n / 2, ids + n, [&](size_t a, size_t b) {
            return leaves[a].bbox.min[axis] + leaves[a].bbox.max[axis] < leaves[b].bbox.min[axis] + leaves[b].bbox.max[axis];
        });
        TreeNode<Real> lhs = median_build(leaves, ids, n / 2);
        TreeNode<Real> rhs = median_build(leaves, ids + n / 2, n - n / 2);
        TreeNode<Real> out;
        out.bbox = surrounding_box(lhs.bbox, rhs.bbox);
        out.leaf = false;
        out.lhs = tree.size(); tree.push_back(lhs);
        out.rhs = tree.size(); tree.push_back(rhs);
        return out;
    }
};

struct TraversalCounters { uint64_t pops = 0, box_hits = 0, leaf_tests = 0; size_t max_stack = 0; };

// bbox_tree.rs:56-91.  Returns leaf index or -1.
template <class Real, class Leaves>
inline long hit_workspace(const BboxTree<Real>& bt, const Leaves& leaves, std::vector<size_t>& stack, const Ray<Real>& ray,
                          Real t_min, Real t_max, HitRecord<Real>& rec, TraversalCounters* ctr) {
    if (!bt.has_root) return -1;
    stack.clear();
    stack.push_back(bt.root);
    long closest = -1;
    while (!stack.empty()) {
        size_t node_idx = stack.back(); stack.pop_back();
        Real t_closest = closest >= 0 ? rec.t : t_max;
        const TreeNode<Real>& node = bt.tree[node_idx];
        if (ctr) ctr->pops++;
        if (!node.bbox.hit2(ray, t_min, t_closest)) continue;
        if (ctr) ctr->box_hits++;
        if (!node.leaf) {
            stack.push_back(node.lhs);
            stack.push_back(node.rhs);
            if (ctr && stack.size() > ctr->max_stack) ctr->max_stack = stack.size();
        } else {
            if (ctr) ctr->leaf_tests++;
            HitRecord<Real> h;
            if (leaves[node.leaf_idx()].hit(ray, t_min, t_closest, h)) { rec = h; closest = (long)node.leaf_idx(); }
        }
    }
    return closest;
}

// ------------------------------------------------------------------------------------------
// material/perlin/mod.rs
// ------------------------------------------------------------------------------------------
template <class Real> struct Perlin {
    Vec3<Real> ranfloat[256];
    int perm_x[256], perm_y[256], perm_z[256];
    static Real interp(const Vec3<Real> kernel[8], Real u, Real v, Real w) {   // perlin/mod.rs:40-63
        Real accum = 0;
        Real uu = u * u * ((Real)3 - (Real)2 * u);
        Real vv = v * v * ((Real)3 - (Real)2 * v);
        Real ww = w * w * ((Real)3 - (Real)2 * w);
        for (int di = 0; di < 2; ++di) { Real i = (Real)di;
            for (int dj = 0; dj < 2; ++dj) { Real j = (Real)dj;
                for (int dk = 0; dk < 2; ++dk) { Real k = (Real)dk;
                    Vec3<Real> weight(u - i, v - j, w - k);
                    accum += (i * uu + ((Real)1 - i) * ((Real)1 - uu)) * (j * vv + ((Real)1 - j) * ((Real)1 - vv)) *
                             (k * ww + ((Real)1 - k) * ((Real)1 - ww)) * kernel[di * 4 + dj * 2 + dk].dot(weight);
                }
            }
        }
        return accum;
    }
    Real noise(const Vec3<Real>& p) const {   // perlin/mod.rs:87-109
        Real xf = std::floor(p.x), yf = std::floor(p.y), zf = std::floor(p.z);
        Real u = p.x - xf, v = p.y - yf, w = p.z - zf;
        // `xf as i32 as usize`, then `(i + di) & 0xFF`: Rust float->int casts saturate.
        auto to_i32 = [](Real f) -> uint32_t {
            if (std::isnan(f)) return 0u;
            if (f >= (Real)2147483647.0) return (uint32_t)2147483647;
            if (f <= (Real)-2147483648.0) return (uint32_t)0x80000000u;
            return (uint32_t)(int32_t)f;
        };
        uint32_t i = to_i32(xf), j = to_i32(yf), k = to_i32(zf);
        Vec3<Real> kernel[8];
        for (uint32_t di = 0; di < 2; ++di)
            for (uint32_t dj = 0; dj < 2; ++dj)
                for (uint32_t dk = 0; dk < 2; ++dk) {
                    int idx = perm_x[(i + di) & 0xFF] ^ perm_y[(j + dj) & 0xFF] ^ perm_z[(k + dk) & 0xFF];
                    kernel[di * 4 + dj * 2 + dk] = ranfloat[idx];
                }
        return interp(kernel, u, v, w);
    }
    Real turbulence(const Vec3<Real>& p, int depth) const {   // perlin/mod.rs:111-123
        Real accum = 0; Vec3<Real> tp = p; Real weight = 1;
        for (int d = 0; d < depth; ++d) { accum += weight * noise(tp); weight *= (Real)0.5; tp = tp.scale((Real)2); }
        return std::fabs(accum);
    }
};

// ------------------------------------------------------------------------------------------
// Scene (scene/mod.rs) with materials (material/*.rs) and textures (material/texture/*.rs)
// ------------------------------------------------------------------------------------------
template <class Real> struct Scatter { Ray<Real> direction; Vec3<Real> attenuation; };   // material/mod.rs:14-18

template <class Real> struct Scene {
    std::vector<GeometricObject<Real>> objects;   // hit id = index (SceneBuilder::add order)
    std::vector<B200rtMaterial> materials;
    std::vector<B200rtTexture> textures;
    struct Image { uint32_t w, h; std::vector<uint8_t> rgb; };
    std::vector<Image> images;
    std::vector<Perlin<Real>> perlin;
    B200rtSkybox skybox{};
    // bounded objects only (scene/mod.rs:124-128); inverted boxes (negative radius) stay in
    // the tree and simply never pass hit2, as in the reference.
    std::vector<long> leaf_to_id;
    BboxTree<Real> tree;

    struct LeafView {
        const Scene* s;
        struct Proxy { const GeometricObject<Real>* g;
            bool hit(const Ray<Real>& r, Real a, Real b, HitRecord<Real>& o) const { return g->hit(r, a, b, o); } };
        Proxy operator[](size_t i) const { return Proxy{&s->objects[(size_t)s->leaf_to_id[i]]}; }
    };

    void load(const B200rtSceneDesc& d, bool reference_topology) {
        objects.resize(d.n_prims);
        for (uint32_t i = 0; i < d.n_prims; ++i) {
            GeometricObject<Real>& g = objects[i];
            g.type = d.prims[i].type;
            uint32_t k = d.prims[i].index;
            if (g.type == B200RT_PRIM_SPHERE) {
                const B200rtSphere& s = d.spheres[k];
                g.sphere.center = {(Real)s.cx, (Real)s.cy, (Real)s.cz}; g.sphere.radius = (Real)s.radius;
            } else if (g.type == B200RT_PRIM_BOX) {
                const B200rtBox& b = d.boxes[k];
                g.box = RectBox<Real>({(Real)b.min[0], (Real)b.min[1], (Real)b.min[2]}, {(Real)b.max[0], (Real)b.max[1], (Real)b.max[2]});
            } else {
                const B200rtRect& r = d.rects[k];
                int d1 = 0, d2 = 1;
                if (r.kind == B200RT_PRIM_RECT_YZ) { d1 = 1; d2 = 2; } else if (r.kind == B200RT_PRIM_RECT_XZ) { d1 = 0; d2 = 2; }
                g.rect = make_rect<Real>(d1, d2, (Real)r.d1_min, (Real)r.d1_max, (Real)r.d2_min, (Real)r.d2_max, (Real)r.offset);
            }
        }
        materials.assign(d.materials, d.materials + d.n_prims);
        textures.assign(d.textures, d.textures + d.n_textures);
        images.resize(d.n_images);
        for (uint32_t i = 0; i < d.n_images; ++i) {
            images[i].w = d.images[i].width; images[i].h = d.images[i].height;
            images[i].rgb.assign(d.images[i].rgb8, d.images[i].rgb8 + (size_t)3 * images[i].w * images[i].h);
        }
        perlin.resize(d.n_perlin);
        for (uint32_t i = 0; i < d.n_perlin; ++i)
            for (int k = 0; k < 256; ++k) {
                perlin[i].ranfloat[k] = {(Real)d.perlin[i].ranfloat[k][0], (Real)d.perlin[i].ranfloat[k][1], (Real)d.perlin[i].ranfloat[k][2]};
                perlin[i].perm_x[k] = d.perlin[i].perm_x[k]; perlin[i].perm_y[k] = d.perlin[i].perm_y[k]; perlin[i].perm_z[k] = d.perlin[i].perm_z[k];
            }
        skybox = d.skybox;
        std::vector<Aabb<Real>> boxes;
        leaf_to_id.clear();
        for (size_t i = 0; i < objects.size(); ++i) { boxes.push_back(objects[i].bounding_box()); leaf_to_id.push_back((long)i); }
        tree.construct(boxes, reference_topology);
    }

    // WorkspaceScene::hit_workspace, scene/mod.rs:153-163 (the unbounded HitList is always
    // empty: every geometry has a bounding box).
    long hit(std::vector<size_t>& stack, const Ray<Real>& ray, Real t_min, Real t_max, HitRecord<Real>& rec, TraversalCounters* ctr) const {
        long leaf = hit_workspace(tree, LeafView{this}, stack, ray, t_min, t_max, rec, ctr);
        return leaf < 0 ? -1 : leaf_to_id[(size_t)leaf];
    }

    // Texture::value dispatch
    Vec3<Real> texture_value(int tex, Real u, Real v, const Vec3<Real>& p) const {
        for (;;) {
            const B200rtTexture& t = textures[(size_t)tex];
            switch (t.kind) {
                case B200RT_TEX_SOLID: return {(Real)t.rgb[0], (Real)t.rgb[1], (Real)t.rgb[2]};   // solid.rs:17-21
                case B200RT_TEX_CHECKER: {                                                          // checker.rs:27-37
                    Real size = (Real)t.scalar;
                    Real sines = std::sin(size * p.x) * std::sin(size * p.y) * std::sin(size * p.z);
                    tex = sines < (Real)0 ? t.odd : t.even;
                    continue;
                }
                case B200RT_TEX_IMAGE: {                                                            // image_texture.rs:34-56
                    const Image& im = images[(size_t)t.image];
                    auto clamp01 = [](Real x) { return x < (Real)0 ? (Real)0 : (x > (Real)1 ? (Real)1 : x); };
                    Real uu = clamp01(u);
                    Real vv = (Real)1 - clamp01(v);
                    auto to_u32 = [](Real f) -> uint32_t { if (!(f > (Real)0)) return 0u; if (f >= (Real)4294967295.0) return 4294967295u; return (uint32_t)f; };
                    uint32_t i = to_u32(uu * (Real)(im.w - 1));
                    uint32_t j = to_u32(vv * (Real)(im.h - 1));
                    const uint8_t* px = &im.rgb[((size_t)j * im.w + i) * 3];
                    Real cs = (Real)1 / (Real)255;
                    return {(Real)px[0] * cs, (Real)px[1] * cs, (Real)px[2] * cs};
                }
                default: {                                                                          // perlin/mod.rs:162-184 (marble)
                    const Perlin<Real>& n = perlin[(size_t)t.image];
                    Real scale = (Real)t.scalar;
                    Real turb = (Real)10 * n.turbulence(p, 7);
                    Vec3<Real> dimm_scale((Real)1 / (Real)5, (Real)1 / (Real)10, (Real)1);
                    Vec3<Real> dimm_weight = Vec3<Real>(0, 0, 1).unit();
                    Vec3<Real> q = dimm_scale.scale(scale) * p;
                    Vec3<Real> vec_dimm(std::sin(q.x + turb), std::sin(q.y + turb), std::sin(q.z + turb));
                    Real total_noise = vec_dimm.dot(dimm_weight);
                    Real noise = (Real)0.5 * ((Real)1 + total_noise);
                    return Vec3<Real>(1, 1, 1).scale(noise);
                }
            }
        }
    }

    // core/math.rs:33-45 random_in_unit_sphere (guess and check)
    template <class U> static Vec3<Real> random_in_unit_sphere(U& rng) {
        for (;;) {
            Real x = (Real)-1 + (Real)2 * rng.template gen<Real>();
            Real y = (Real)-1 + (Real)2 * rng.template gen<Real>();
            Real z = (Real)-1 + (Real)2 * rng.template gen<Real>();
            Vec3<Real> p(x, y, z);
            if (p.length_squared() <= (Real)1) return p;
        }
    }
    template <class U> static Vec3<Real> random_unit_vector(U& rng) { return random_in_unit_sphere(rng).unit(); }   // math.rs:62-68
    template <class U> static Vec3<Real> random_in_unit_disk(U& rng) {                                             // math.rs:70-81
        for (;;) {
            Real x = (Real)-1 + (Real)2 * rng.template gen<Real>();
            Real y = (Real)-1 + (Real)2 * rng.template gen<Real>();
            Vec3<Real> p(x, y, 0);
            if (p.length_squared() <= (Real)1) return p;
        }
    }

    static Real reflectance(Real cosine, Real ref_idx) {   // dielectric.rs:15-19
        Real r0 = ((Real)1 - ref_idx) / ((Real)1 + ref_idx);
        r0 = r0 * r0;
        return r0 + ((Real)1 - r0) * std::pow((Real)1 - cosine, (Real)5);
    }

    // MaterialType::scatter, material_type.rs:50-64
    template <class U> bool scatter(long id, U& rng, const Ray<Real>& ray, const HitRecord<Real>& rec, Scatter<Real>& out) const {
        const B200rtMaterial& m = materials[(size_t)id];
        switch (m.kind) {
            case B200RT_MAT_METAL: {   // metal.rs:27-39 — always Some, sampler drawn even for fuzz 0
                Vec3<Real> reflected = ray.direction.unit().reflect(rec.normal);
                Vec3<Real> dir = reflected + random_in_unit_sphere(rng).scale((Real)m.param);
                out.direction = {rec.point, dir};
                out.attenuation = {(Real)m.albedo[0], (Real)m.albedo[1], (Real)m.albedo[2]};
                return true;
            }
            case B200RT_MAT_DIELECTRIC: {   // dielectric.rs:22-49
                Real ir = (Real)m.param;
                Real refraction_ratio = rec.front_face ? (Real)1 / ir : ir;
                Vec3<Real> unit_direction = ray.direction.unit();
                Real cos_theta = fmin_<Real>(unit_direction.scale((Real)-1).dot(rec.normal), (Real)1);
                Real sin_theta = std::sqrt((Real)1 - cos_theta * cos_theta);
                Vec3<Real> direction;
                if (refraction_ratio * sin_theta > (Real)1 || reflectance(cos_theta, refraction_ratio) > rng.template gen<Real>())
                    direction = unit_direction.reflect(rec.normal);
                else
                    direction = unit_direction.refract(rec.normal, refraction_ratio);
                out.direction = {rec.point, direction};
                out.attenuation = {1, 1, 1};
                return true;
            }
            case B200RT_MAT_LAMBERTIAN:     // lambertian.rs:22-36
            case B200RT_MAT_FAIRY_LIGHT: {  // lighting.rs:43-57
                Vec3<Real> sc = rec.normal + random_unit_vector(rng);
                if (sc.near_zero()) sc = rec.normal;
                out.direction = {rec.point, sc};
                Vec3<Real> a = texture_value(m.texture, rec.u, rec.v, rec.point);
                out.attenuation = m.kind == B200RT_MAT_FAIRY_LIGHT ? a.unit() : a;
                return true;
            }
            default: return false;          // DiffuseLight, lighting.rs:26-28
        }
    }
    // MaterialType::emitted, material_type.rs:66-78
    bool emitted(long id, const Ray<Real>& ray, const HitRecord<Real>& rec, Vec3<Real>& out) const {
        const B200rtMaterial& m = materials[(size_t)id];
        if (m.kind == B200RT_MAT_DIFFUSE_LIGHT) { out = texture_value(m.texture, rec.u, rec.v, rec.point); return true; }   // lighting.rs:21-24
        if (m.kind == B200RT_MAT_FAIRY_LIGHT) {                                                                              // lighting.rs:59-66
            Vec3<Real> src = texture_value(m.texture, rec.u, rec.v, rec.point);
            Real scale = rec.normal.dot(ray.direction.scale((Real)-1));
            out = src.scale(scale / ray.direction.length());
            return true;
        }
        return false;
    }
    Vec3<Real> background(const Ray<Real>& r) const {   // skybox/mod.rs:5-25
        if (skybox.kind == B200RT_SKY_ABOVE) {
            Vec3<Real> unit = r.direction.unit();
            Real t = (Real)0.5 * (unit.y + (Real)1);
            return Vec3<Real>(1, 1, 1).scale((Real)1 - t) + Vec3<Real>((Real)0.5, (Real)0.7, (Real)1.0).scale(t);
        }
        if (skybox.kind == B200RT_SKY_FLAT) return {(Real)skybox.rgb[0], (Real)skybox.rgb[1], (Real)skybox.rgb[2]};
        return {0, 0, 0};
    }
};

// ------------------------------------------------------------------------------------------
// camera/mod.rs:98-131
// ------------------------------------------------------------------------------------------
template <class Real> struct Camera {
    Real height, width, focal_length, lens_radius; bool has_lens;
    uint32_t W, H;
    Vec3<Real> origin, w, u, v; Real focus_length;
    explicit Camera(const B200rtCamera& c) {
        height = (Real)c.height; width = (Real)c.width; focal_length = (Real)c.focal_length;
        has_lens = c.lens_radius >= 0; lens_radius = (Real)c.lens_radius;
        W = c.image_width; H = c.image_height;
        origin = {(Real)c.origin[0], (Real)c.origin[1], (Real)c.origin[2]};
        w = {(Real)c.w[0], (Real)c.w[1], (Real)c.w[2]};
        u = {(Real)c.u[0], (Real)c.u[1], (Real)c.u[2]};
        v = {(Real)c.v[0], (Real)c.v[1], (Real)c.v[2]};
        focus_length = (Real)c.focus_length;
    }
    template <class U> Ray<Real> pixel_ray(U& rng, Real x, Real y) const {
        Real x_percent = x / (Real)W;
        Real y_percent = y / (Real)H;
        Vec3<Real> horizontal = u.scale(width * focus_length);
        Vec3<Real> vertical = v.scale(height * focus_length);
        Vec3<Real> lower_left = origin - horizontal.scale((Real)0.5) - vertical.scale((Real)0.5) - w.scale(focal_length * focus_length);
        Vec3<Real> offset;
        if (has_lens) {
            Vec3<Real> rd = Scene<Real>::random_in_unit_disk(rng).scale(lens_radius);
            offset = u.scale(rd.x) + v.scale(rd.y);
        }
        Vec3<Real> direction = lower_left + horizontal.scale(x_percent) + vertical.scale(y_percent) - origin - offset;
        return {origin + offset, direction};
    }
};

// ------------------------------------------------------------------------------------------
// render.rs:17-70
// ------------------------------------------------------------------------------------------
struct RenderCounters { uint64_t rays = 0, paths = 0, depth_exhausted = 0; TraversalCounters trav; };

template <class Real, class U>
inline Vec3<Real> ray_color(U& rng, std::vector<size_t>& stack, const Ray<Real>& incoming, const Scene<Real>& scene, size_t max_depth, RenderCounters* ctr) {
    Ray<Real> ray = incoming;
    Vec3<Real> attenuation(1, 1, 1), emitted(0, 0, 0);
    while (max_depth > 0) {
        HitRecord<Real> r;
        if (ctr) ctr->rays++;
        long id = scene.hit(stack, ray, (Real)0.001, std::numeric_limits<Real>::infinity(), r, ctr ? &ctr->trav : nullptr);
        if (id >= 0) {
            Vec3<Real> e;
            if (scene.emitted(id, ray, r, e)) emitted = emitted + attenuation * e;
            Scatter<Real> sc;
            if (scene.scatter(id, rng, ray, r, sc)) {
                attenuation = attenuation * sc.attenuation;
                ray = sc.direction;
            } else break;
        } else {
            emitted = emitted + attenuation * scene.background(ray);
            break;
        }
        max_depth -= 1;
        if (max_depth == 0 && ctr) ctr->depth_exhausted++;
    }
    return emitted;
}

// render_scanline, render.rs:49-70.  `buf` receives the SUM over samples.  The RNG is keyed
// per (pixel, sample) like the device kernel, instead of one ThreadRng per rayon worker.
template <class Real>
inline void render_scanline(const Camera<Real>& cam, const Scene<Real>& scene, uint64_t seed, uint32_t sample_offset, size_t samples,
                            size_t max_depth, std::vector<size_t>& stack, size_t line_idx, Vec3<Real>* buf, RenderCounters* ctr) {
    for (size_t idx = 0; idx < cam.W; ++idx) {
        Vec3<Real> c(0, 0, 0);
        for (size_t s = 0; s < samples; ++s) {
            UniformSource rng;
            rng.rng = Rng(seed, (uint32_t)(line_idx * cam.W + idx), sample_offset + (uint32_t)s);
            Real jitter_idx = (Real)idx + rng.template gen<Real>();
            Real jitter_line_idx = (Real)line_idx + rng.template gen<Real>();
            Ray<Real> r = cam.pixel_ray(rng, jitter_idx, jitter_line_idx);
            if (ctr) ctr->paths++;
            c = c + ray_color<Real>(rng, stack, r, scene, max_depth, ctr);
        }
        buf[idx] = c;
    }
}

// image.rs:31-44 + core/color.rs:31-38: mean, sqrt gamma, saturating u8 cast, vertical flip.
inline uint8_t sat_u8(double x) {   // Rust `as u8`: saturates, NaN -> 0
    if (!(x > 0.0)) return 0;
    if (x >= 255.0) return 255;
    return (uint8_t)x;
}
template <class Real>
inline void to_image(const Vec3<Real>* data, uint32_t W, uint32_t H, size_t samples, uint8_t* out_rgb8) {
    for (uint32_t j = 0; j < H; ++j)
        for (uint32_t i = 0; i < W; ++i) {
            Vec3<Real> c = data[(size_t)j * W + i].scale((Real)1 / (Real)samples);
            c = {std::sqrt(c.x), std::sqrt(c.y), std::sqrt(c.z)};
            uint8_t* px = out_rgb8 + ((size_t)(H - j - 1) * W + i) * 3;
            px[0] = sat_u8((double)(c.x * (Real)255.999));
            px[1] = sat_u8((double)(c.y * (Real)255.999));
            px[2] = sat_u8((double)(c.z * (Real)255.999));
        }
}

}  // namespace oracle
