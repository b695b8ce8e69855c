// rt_device.cuh — device-side data layout, math, RNG, intersection, traversal and shading
// for the sm_100a path tracer.  Everything here is f32.  Reference citations are relative
// to /root/reference/src/raytracer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rt.h"

namespace b200rt {

// ------------------------------------------------------------------------------------------
// Device scene layout (all arrays 16-byte aligned, read-only during a launch)
// ------------------------------------------------------------------------------------------
// BVH2 inner node, 64 B = 4 x float4 (one 16-byte vector load each):
//   q0 = (c0.lo.x, c0.lo.y, c0.lo.z, c0.hi.x)
//   q1 = (c0.hi.y, c0.hi.z, c1.lo.x, c1.lo.y)
//   q2 = (c1.lo.z, c1.hi.x, c1.hi.y, c1.hi.z)
//   q3 = (bits(child0), bits(child1), 0, 0)
// child >= 0: inner-node index; child < 0: leaf, ~child = (prim_type << 28) | hit_id.
// The boxes of BOTH children live in the parent, so one node fetch feeds two slab tests
// (the reference's TreeNode is 72 B of f64 per box and tests at pop, bvh/bbox_tree.rs:16-20,
// :73-76).  Nodes are stored breadth-first: the first K nodes are the top of the tree and
// are the ones staged in shared memory when the whole tree does not fit.
struct BvhNode { float4 q0, q1, q2, q3; };

#define B200RT_LEAF_TYPE_SHIFT 28
#define B200RT_LEAF_ID_MASK 0x0FFFFFFFu
#define B200RT_EMPTY_LEAF ((int)0x80000000)   // ~0x7FFFFFFF: never produced for a real prim

// Geometry record per hit id, 32 B = 2 x float4:
//   sphere: g0 = (cx, cy, cz, r)
//   rect:   g0 = (d1_min, d1_max, d2_min, d2_max), g1.x = offset
//   box:    g0 = (min.xyz, -), g1 = (max.xyz, -)
struct GeomRec { float4 g0, g1; };

// Material record per hit id, 32 B:
//   m0 = (r, g, b, param)   Metal: albedo + fuzz; Dielectric: param = ir;
//                           Lambertian/lights with a SOLID texture: rgb resolved at upload
//   kind = B200RT_MAT_*; tex = texture index or -1 when rgb is already resolved
struct MatRec { float4 m0; uint32_t kind; int32_t tex; uint32_t pad0, pad1; };

struct TexRec { uint32_t kind; float r, g, b; float scalar; int32_t odd, even, image; };   // 32 B

struct ImageRec { const uchar4* texels; uint32_t width, height; uint32_t pad; };
struct PerlinRec { float4 ranfloat[256]; uint8_t perm_x[256], perm_y[256], perm_z[256]; };

struct DeviceScene {
    const BvhNode* nodes;
    const GeomRec* geom;
    const MatRec* mats;
    const TexRec* tex;
    const ImageRec* images;
    const PerlinRec* perlin;
    uint32_t n_nodes, n_prims, n_tex, bvh_depth;
    uint32_t sky_kind; float sky_r, sky_g, sky_b;
};

// Camera constants, reduced on the host in f64 from camera/mod.rs:98-114:
//   dir = llo + hw * x + vh * y - offset,   llo = lower_left - origin, hw = horizontal / W,
//   vh = vertical / H;  offset = ul * rd.x + vl * rd.y with ul = u * lens_radius.
struct DeviceCamera {
    float3 origin, llo, hw, vh, ul, vl;
    uint32_t width, height;
    int has_lens;
};


// ------------------------------------------------------------------------------------------
// Scene accessors.  The same traversal/shading code runs over a scene staged entirely in
// shared memory (SmemAcc: Weekend-sized scenes, ~60 KB, divergent 16-byte LDS cost 4
// wavefronts per warp instead of up to 32 L1 tag lookups) or resident in global memory
// behind the read-only path (GmemAcc), with the top `n_top` breadth-first BVH nodes still
// in shared memory (hot top levels).
// ------------------------------------------------------------------------------------------
struct SmemAcc {
    const float4* nodes; const float4* geom; const float4* mats; const float4* tex;
    __device__ __forceinline__ float4 node_q(int node, int k) const { return nodes[node * 4 + k]; }
    __device__ __forceinline__ float4 geom0(int id) const { return geom[id * 2]; }
    __device__ __forceinline__ float4 geom1(int id) const { return geom[id * 2 + 1]; }
    __device__ __forceinline__ MatRec mat(int id) const {
        float4 a = mats[id * 2], b = mats[id * 2 + 1];
        MatRec m; m.m0 = a; m.kind = __float_as_uint(b.x); m.tex = __float_as_int(b.y); m.pad0 = m.pad1 = 0; return m;
    }
    __device__ __forceinline__ TexRec texrec(int t) const {
        float4 a = tex[t * 2], b = tex[t * 2 + 1];
        TexRec T; T.kind = __float_as_uint(a.x); T.r = a.y; T.g = a.z; T.b = a.w;
        T.scalar = b.x; T.odd = __float_as_int(b.y); T.even = __float_as_int(b.z); T.image = __float_as_int(b.w); return T;
    }
};
struct GmemAcc {
    const float4* __restrict__ nodes; const float4* __restrict__ geom; const float4* __restrict__ mats; const float4* __restrict__ tex;
    const float4* top; int n_top;   // shared-memory copy of nodes [0, n_top)
    __device__ __forceinline__ float4 node_q(int node, int k) const {
        return node < n_top ? top[node * 4 + k] : __ldg(nodes + node * 4 + k);
    }
    __device__ __forceinline__ float4 geom0(int id) const { return __ldg(geom + id * 2); }
    __device__ __forceinline__ float4 geom1(int id) const { return __ldg(geom + id * 2 + 1); }
    __device__ __forceinline__ MatRec mat(int id) const {
        float4 a = __ldg(mats + id * 2), b = __ldg(mats + id * 2 + 1);
        MatRec m; m.m0 = a; m.kind = __float_as_uint(b.x); m.tex = __float_as_int(b.y); m.pad0 = m.pad1 = 0; return m;
    }
    __device__ __forceinline__ TexRec texrec(int t) const {
        float4 a = __ldg(tex + t * 2), b = __ldg(tex + t * 2 + 1);
        TexRec T; T.kind = __float_as_uint(a.x); T.r = a.y; T.g = a.z; T.b = a.w;
        T.scalar = b.x; T.odd = __float_as_int(b.y); T.even = __float_as_int(b.z); T.image = __float_as_int(b.w); return T;
    }
};

// ------------------------------------------------------------------------------------------
// float3 helpers.  Plain operators: nvcc may contract a*b+c into FMA here.  The
// intersection routines that must be reproducible bit-for-bit by oracle/gpu_f32.hpp use
// explicit __f*_rn / fmaf instead.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 fma3(float s, float3 a, float3 b) { return f3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ float comp(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ------------------------------------------------------------------------------------------
// RNG: counter-keyed PCG (documented in include/b200rt.h).  Replaces rand::ThreadRng
// (src/main.rs:98,119): a (pixel, sample) pair owns its stream, so an image does not
// depend on how tiles or sample ranges are sharded over warps or GPUs.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t pcg_out(uint32_t s) {
    uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) { return pcg_out(x * 747796405u + 2891336453u); }

struct RngKeys { uint32_t k0, k1; };
__host__ __device__ __forceinline__ RngKeys rng_keys(uint64_t seed) {
    RngKeys k;
    k.k0 = hash32((uint32_t)seed);
    k.k1 = hash32((uint32_t)(seed >> 32) ^ k.k0);
    return k;
}

struct Rng {
    uint32_t state, inc;
    __device__ __forceinline__ void init(RngKeys k, uint32_t a, uint32_t b) {
        state = hash32(b + hash32(a ^ k.k1));
        inc = (hash32(a + hash32(b ^ k.k0)) << 1) | 1u;
    }
    __device__ __forceinline__ uint32_t next_u32() {
        state = state * 747796405u + inc;
        return pcg_out(state);
    }
    // rng.gen::<f64>() (core/math.rs:24) at 24-bit resolution
    __device__ __forceinline__ float gen() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    // random_real(rng, -1, 1) = min + (max - min) * gen  (core/math.rs:23-25)
    __device__ __forceinline__ float gen_pm1() { return fmaf(2.0f, gen(), -1.0f); }
};

// ------------------------------------------------------------------------------------------
// Ray with the per-ray constants hoisted out of Aabb::hit2 (bvh/aabb.rs:66: `1.0 / d` is
// recomputed per node per axis in the reference) and Sphere::hit (sphere.rs:31: a = d.d).
// ------------------------------------------------------------------------------------------
struct RayF {
    float3 o, d, inv;   // inv = 1/d (IEEE, may be +-inf)
    float a, inv_a;     // d.d and 1/(d.d)
};
__device__ __forceinline__ RayF make_ray(float3 o, float3 d) {
    RayF r;
    r.o = o; r.d = d;
    r.inv = f3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));
    r.a = fmaf(d.z, d.z, fmaf(d.y, d.y, __fmul_rn(d.x, d.x)));
    r.inv_a = __frcp_rn(r.a);
    return r;
}

// Aabb::hit2 (bvh/aabb.rs:62-79) in f32 with the reference's exact NaN/inf behaviour:
// the swap is selected by the sign of 1/d (not by min/max of the two plane distances), a
// NaN plane distance (0 * inf) leaves the interval unchanged, `t_max <= t_min` misses.
// Returns the entry distance through *t_enter.
__device__ __forceinline__ bool aabb_hit2(const RayF& r, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                          float t_min, float t_max, float* t_enter) {
    float ax = __fmul_rn(__fsub_rn(lox, r.o.x), r.inv.x), bx = __fmul_rn(__fsub_rn(hix, r.o.x), r.inv.x);
    float ay = __fmul_rn(__fsub_rn(loy, r.o.y), r.inv.y), by = __fmul_rn(__fsub_rn(hiy, r.o.y), r.inv.y);
    float az = __fmul_rn(__fsub_rn(loz, r.o.z), r.inv.z), bz = __fmul_rn(__fsub_rn(hiz, r.o.z), r.inv.z);
    bool sx = r.inv.x < 0.0f, sy = r.inv.y < 0.0f, sz = r.inv.z < 0.0f;
    float t0x = sx ? bx : ax, t1x = sx ? ax : bx;
    float t0y = sy ? by : ay, t1y = sy ? ay : by;
    float t0z = sz ? bz : az, t1z = sz ? az : bz;
    // `if t0 > t_min {t0} else {t_min}` keeps t_min when t0 is NaN == fmaxf semantics.
    // The reference exits after each axis; the interval only ever shrinks, so testing once
    // after all three axes gives the same answer.
    float lo = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, t_min));
    float hi = fminf(fminf(t1x, t1y), fminf(t1z, t_max));
    *t_enter = lo;
    return !(hi <= lo);
}

// ------------------------------------------------------------------------------------------
// Primitive intersection.  Closest-hit bookkeeping: `best_t` shrinks; the interval is
// inclusive at both ends (sphere.rs:41-46, rect.rs:58-60); among exactly equal t the
// highest hit id wins (what HitList's linear scan does, scene/mod.rs:66-76).
// These use explicit-rounding intrinsics only, so oracle/gpu_f32.hpp reproduces them
// bit-for-bit on the CPU.
// ------------------------------------------------------------------------------------------
struct Closest {
    float t;       // current closest distance (starts at t_max)
    int code;      // (type << 28) | id of the closest prim, -1 = none
    int face;      // box only: which of the 6 rects (rect.rs:149-154 order) was hit
};

__device__ __forceinline__ bool accept_t(float t, float t_min, const Closest& c, int id) {
    // NaN-safe: a NaN root is rejected.
    return (t >= t_min) && (t < c.t || (t == c.t && (c.code < 0 || id > (int)(c.code & B200RT_LEAF_ID_MASK))));
}

// Sphere::hit (geometry/sphere.rs:29-52).  Same roots as the reference's textbook
// quadratic, computed in the cancellation-free form (the discriminant from the closest-
// approach vector, the roots as c/q and q/a) so that f32 stays inside 1e-5 of the f64
// reference.
__device__ __forceinline__ bool sphere_roots(const RayF& r, float4 s, float* root_lo, float* root_hi) {
    float ocx = __fsub_rn(r.o.x, s.x), ocy = __fsub_rn(r.o.y, s.y), ocz = __fsub_rn(r.o.z, s.z);
    float bp = -fmaf(ocz, r.d.z, fmaf(ocy, r.d.y, __fmul_rn(ocx, r.d.x)));   // -half_b
    float q = __fmul_rn(bp, r.inv_a);
    float lx = fmaf(q, r.d.x, ocx), ly = fmaf(q, r.d.y, ocy), lz = fmaf(q, r.d.z, ocz);
    float l2 = fmaf(lz, lz, fmaf(ly, ly, __fmul_rn(lx, lx)));
    float r2 = __fmul_rn(s.w, s.w);
    float delta = __fsub_rn(r2, l2);               // discriminant / a
    if (delta < 0.0f) return false;                // sphere.rs:35-37
    float sq = __fsqrt_rn(__fmul_rn(delta, r.a));  // sqrt(discriminant)
    float qq = __fadd_rn(bp, copysignf(sq, bp));
    float c = __fsub_rn(fmaf(ocz, ocz, fmaf(ocy, ocy, __fmul_rn(ocx, ocx))), r2);
    float r0 = __fdiv_rn(c, qq);
    float r1 = __fmul_rn(qq, r.inv_a);
    *root_lo = fminf(r0, r1);
    *root_hi = fmaxf(r0, r1);
    return true;
}

__device__ __forceinline__ void hit_sphere(const RayF& r, float4 s, int id, float t_min, Closest& c) {
    float lo, hi;
    if (!sphere_roots(r, s, &lo, &hi)) return;
    // sphere.rs:41-46: try the near root, then the far root, against [t_min, closest]
    float root = lo;
    bool ok = accept_t(root, t_min, c, id);
    if (!ok) { root = hi; ok = accept_t(root, t_min, c, id); }
    if (ok) { c.t = root; c.code = (int)((B200RT_PRIM_SPHERE << B200RT_LEAF_TYPE_SHIFT) | (uint32_t)id); }
}

// Rect<D1,D2>::hit (geometry/rect.rs:55-80) for the plane `offset` on axis dn with in-plane
// axes d1, d2.  Returns t or NaN.
__device__ __forceinline__ float rect_t(const RayF& r, int d1, int d2, float d1_min, float d1_max, float d2_min, float d2_max,
                                        float offset, float t_min, float t_max) {
    int dn = 3 - d1 - d2;
    float t = __fmul_rn(__fsub_rn(offset, comp(r.o, dn)), comp(r.inv, dn));
    if (!(t >= t_min && t <= t_max)) return __int_as_float(0x7fc00000);
    float a = fmaf(t, comp(r.d, d1), comp(r.o, d1));
    float b = fmaf(t, comp(r.d, d2), comp(r.o, d2));
    if (a < d1_min || a > d1_max || b < d2_min || b > d2_max) return __int_as_float(0x7fc00000);
    return t;
}

__device__ __forceinline__ void rect_axes(uint32_t type, int& d1, int& d2) {
    d1 = (type == B200RT_PRIM_RECT_YZ) ? 1 : 0;
    d2 = (type == B200RT_PRIM_RECT_XY) ? 1 : 2;
}

__device__ __forceinline__ void hit_rect(const RayF& r, uint32_t type, float4 g0, float offset, int id, float t_min, Closest& c) {
    int d1, d2;
    rect_axes(type, d1, d2);
    float t = rect_t(r, d1, d2, g0.x, g0.y, g0.z, g0.w, offset, t_min, c.t);
    if (accept_t(t, t_min, c, id)) { c.t = t; c.code = (int)((type << B200RT_LEAF_TYPE_SHIFT) | (uint32_t)id); }
}

// RectBox::hit (geometry/rect.rs:147-156): the six faces in the reference's order, each
// against the shrinking interval; a later face replaces an earlier one at equal t.
__device__ __forceinline__ void hit_box(const RayF& r, float4 lo, float4 hi, int id, float t_min, Closest& c) {
    float best = c.t;
    int face = -1;
    // rect.rs:116-127: xy(z = max), xy(z = min), yz(x = max), yz(x = min), xz(y = max), xz(y = min)
    float t;
    t = rect_t(r, 0, 1, lo.x, hi.x, lo.y, hi.y, hi.z, t_min, best); if (t <= best) { best = t; face = 0; }
    t = rect_t(r, 0, 1, lo.x, hi.x, lo.y, hi.y, lo.z, t_min, best); if (t <= best) { best = t; face = 1; }
    t = rect_t(r, 1, 2, lo.y, hi.y, lo.z, hi.z, hi.x, t_min, best); if (t <= best) { best = t; face = 2; }
    t = rect_t(r, 1, 2, lo.y, hi.y, lo.z, hi.z, lo.x, t_min, best); if (t <= best) { best = t; face = 3; }
    t = rect_t(r, 0, 2, lo.x, hi.x, lo.z, hi.z, hi.y, t_min, best); if (t <= best) { best = t; face = 4; }
    t = rect_t(r, 0, 2, lo.x, hi.x, lo.z, hi.z, lo.y, t_min, best); if (t <= best) { best = t; face = 5; }
    if (face >= 0 && accept_t(best, t_min, c, id)) {
        c.t = best; c.face = face;
        c.code = (int)((B200RT_PRIM_BOX << B200RT_LEAF_TYPE_SHIFT) | (uint32_t)id);
    }
}

// GeometricObject::hit dispatch (geometry/object.rs:44-58)
template <class Acc>
__device__ __forceinline__ void hit_leaf(const RayF& r, const Acc& acc, int leaf, float t_min, Closest& c) {
    uint32_t code = (uint32_t)~leaf;
    uint32_t type = code >> B200RT_LEAF_TYPE_SHIFT;
    int id = (int)(code & B200RT_LEAF_ID_MASK);
    if (type > B200RT_PRIM_BOX) return;   // B200RT_EMPTY_LEAF
    float4 g0 = acc.geom0(id);
    if (type == B200RT_PRIM_SPHERE) {
        hit_sphere(r, g0, id, t_min, c);
    } else {
        float4 g1 = acc.geom1(id);
        if (type == B200RT_PRIM_BOX) hit_box(r, g0, g1, id, t_min, c);
        else hit_rect(r, type, g0, g1.x, id, t_min, c);
    }
}

// ------------------------------------------------------------------------------------------
// BVH closest-hit.  Replaces BboxTree::hit_workspace (bvh/bbox_tree.rs:56-91): instead of
// popping every child and testing its box at pop (children pushed untested, rhs first, no
// ordering), one node fetch tests both children, descends into the nearer one and pushes
// the farther one; `c.t` shrinks exactly like `t_closest`.  Closest-hit results do not
// depend on the visiting order (ties are resolved by id in accept_t).
// `stack` is a per-thread column of a shared-memory array: element k lives at
// stack[k * stride].
// ------------------------------------------------------------------------------------------
struct TravCounters { uint32_t nodes, prims; };

template <bool COUNT, class Acc>
__device__ __forceinline__ void closest_hit(const RayF& r, const Acc& acc, int* stack, int stride,
                                            float t_min, Closest& c, TravCounters& tc) {
    int sp = 0;
    int node = 0;
    for (;;) {
        if (node >= 0) {
            float4 q0 = acc.node_q(node, 0), q1 = acc.node_q(node, 1), q2 = acc.node_q(node, 2), q3 = acc.node_q(node, 3);
            if (COUNT) tc.nodes++;
            float e0, e1;
            bool h0 = aabb_hit2(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, t_min, c.t, &e0);
            bool h1 = aabb_hit2(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, t_min, c.t, &e1);
            int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            if (h0 && h1) {
                bool swap = e1 < e0;
                int near_c = swap ? c1 : c0, far_c = swap ? c0 : c1;
                stack[sp * stride] = far_c; ++sp;
                node = near_c;
            } else if (h0) node = c0;
            else if (h1) node = c1;
            else {
                if (sp == 0) break;
                --sp; node = stack[sp * stride];
            }
        } else {
            if (COUNT) tc.prims++;
            hit_leaf(r, acc, node, t_min, c);
            if (sp == 0) break;
            --sp; node = stack[sp * stride];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Hit record (geometry/hittable.rs:17-37), built once for the winning primitive.
// ------------------------------------------------------------------------------------------
struct HitRec {
    float3 p, n;        // n is face-flipped
    float3 n_out;       // geometric (outward) normal before the flip: Sphere::get_uv input
    float t;
    bool front;
    uint32_t type; int id; int face;
    bool has_uv; float uv_u, uv_v;   // parity hooks inject the caller's (u, v); the render path computes them lazily
};

template <class Acc>
__device__ __forceinline__ HitRec make_hit(const RayF& r, const Acc& acc, const Closest& c) {
    HitRec h;
    h.has_uv = false; h.uv_u = 0.f; h.uv_v = 0.f;
    uint32_t code = (uint32_t)c.code;
    h.type = code >> B200RT_LEAF_TYPE_SHIFT;
    h.id = (int)(code & B200RT_LEAF_ID_MASK);
    h.face = c.face;
    h.t = c.t;
    h.p = fma3(c.t, r.d, r.o);                            // Ray::at, vec3.rs:253
    if (h.type == B200RT_PRIM_SPHERE) {
        float4 s = acc.geom0(h.id);
        float inv_r = __frcp_rn(s.w);                     // sphere.rs:49 `scale(1.0 / radius)`
        // (p - center) / r, evaluated as (oc + t d) / r: one rounding instead of two
        float3 oc = f3(__fsub_rn(r.o.x, s.x), __fsub_rn(r.o.y, s.y), __fsub_rn(r.o.z, s.z));
        float3 hp = fma3(c.t, r.d, oc);
        h.n_out = f3(__fmul_rn(hp.x, inv_r), __fmul_rn(hp.y, inv_r), __fmul_rn(hp.z, inv_r));
    } else {
        int dn;
        if (h.type == B200RT_PRIM_BOX) dn = (c.face < 2) ? 2 : (c.face < 4 ? 0 : 1);
        else { int d1, d2; rect_axes(h.type, d1, d2); dn = 3 - d1 - d2; }
        h.n_out = f3(dn == 0 ? 1.0f : 0.0f, dn == 1 ? 1.0f : 0.0f, dn == 2 ? 1.0f : 0.0f);   // rect.rs:75-76
    }
    float dn_ = fmaf(r.d.z, h.n_out.z, fmaf(r.d.y, h.n_out.y, __fmul_rn(r.d.x, h.n_out.x)));
    h.front = dn_ < 0.0f;                                 // hittable.rs:25
    h.n = h.front ? h.n_out : -h.n_out;
    return h;
}

// u,v of the hit (sphere.rs:18-25, rect.rs:71-72) — only image textures read them, so they
// are computed on demand.
template <class Acc>
__device__ __forceinline__ void hit_uv(const HitRec& h, const Acc& acc, float* u, float* v) {
    const float PI = 3.14159265358979323846f;
    if (h.has_uv) { *u = h.uv_u; *v = h.uv_v; return; }
    if (h.type == B200RT_PRIM_SPHERE) {
        float ny = fminf(fmaxf(-h.n_out.y, -1.0f), 1.0f);   // f32 rounding can leave |n.y| a hair above 1
        float theta = acosf(ny);
        float phi = atan2f(-h.n_out.z, h.n_out.x) + PI;
        *u = phi / (2.0f * PI);
        *v = theta / PI;
    } else {
        float4 g0; int d1, d2;
        if (h.type == B200RT_PRIM_BOX) {
            float4 lo = acc.geom0(h.id), hi = acc.geom1(h.id);
            if (h.face < 2) { d1 = 0; d2 = 1; g0 = make_float4(lo.x, hi.x, lo.y, hi.y); }
            else if (h.face < 4) { d1 = 1; d2 = 2; g0 = make_float4(lo.y, hi.y, lo.z, hi.z); }
            else { d1 = 0; d2 = 2; g0 = make_float4(lo.x, hi.x, lo.z, hi.z); }
        } else { rect_axes(h.type, d1, d2); g0 = acc.geom0(h.id); }
        float a = comp(h.p, d1), b = comp(h.p, d2);
        *u = (a - g0.x) / (g0.y - g0.x);
        *v = (b - g0.z) / (g0.w - g0.z);
    }
}

// ------------------------------------------------------------------------------------------
// Textures (material/texture/*.rs, material/perlin/mod.rs)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float perlin_noise(const PerlinRec* __restrict__ P, float3 p) {   // perlin/mod.rs:87-109
    float xf = floorf(p.x), yf = floorf(p.y), zf = floorf(p.z);
    float u = p.x - xf, v = p.y - yf, w = p.z - zf;
    int i = __float2int_rz(xf), j = __float2int_rz(yf), k = __float2int_rz(zf);   // `as i32` saturates like cvt.rzi
    float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);   // perlin/mod.rs:43-45
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                int idx = P->perm_x[(i + di) & 255] ^ P->perm_y[(j + dj) & 255] ^ P->perm_z[(k + dk) & 255];
                float4 g = P->ranfloat[idx];
                float wx = u - (float)di, wy = v - (float)dj, wz = w - (float)dk;
                float blend = (di ? uu : 1.0f - uu) * (dj ? vv : 1.0f - vv) * (dk ? ww : 1.0f - ww);
                accum += blend * (g.x * wx + g.y * wy + g.z * wz);
            }
    return accum;
}

__device__ __forceinline__ float perlin_turbulence(const PerlinRec* __restrict__ P, float3 p) {   // perlin/mod.rs:111-123, depth 7
    float accum = 0.0f, weight = 1.0f;
#pragma unroll 1
    for (int d = 0; d < 7; ++d) {
        accum += weight * perlin_noise(P, p);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(accum);
}

// Texture::value.  A checker picks exactly one child per level (checker.rs:27-37), so the
// recursion is a loop.
template <class Acc>
__device__ __forceinline__ float3 texture_value(const Acc& acc, const ImageRec* __restrict__ images, const PerlinRec* __restrict__ perlin,
                                                int t, const HitRec& h) {
    for (;;) {
        TexRec T = acc.texrec(t);
        if (T.kind == B200RT_TEX_SOLID) return f3(T.r, T.g, T.b);                       // solid.rs:17-21
        if (T.kind == B200RT_TEX_CHECKER) {                                              // checker.rs:28-36
            // accurate sinf: sizes reach 8/r ~ 160 and coordinates ~ 30
            float s = sinf(T.scalar * h.p.x) * sinf(T.scalar * h.p.y) * sinf(T.scalar * h.p.z);
            t = s < 0.0f ? T.odd : T.even;
            continue;
        }
        if (T.kind == B200RT_TEX_IMAGE) {                                                // image_texture.rs:34-56
            float u, v;
            hit_uv(h, acc, &u, &v);
            ImageRec im = images[T.image];
            u = fminf(fmaxf(u, 0.0f), 1.0f);
            v = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
            uint32_t i = (uint32_t)(u * (float)(im.width - 1));     // truncation: nearest texel below, no filtering
            uint32_t j = (uint32_t)(v * (float)(im.height - 1));
            i = min(i, im.width - 1); j = min(j, im.height - 1);
            uchar4 px = __ldg(&im.texels[(size_t)j * im.width + i]);
            const float cs = 1.0f / 255.0f;
            return f3((float)px.x * cs, (float)px.y * cs, (float)px.z * cs);
        }
        // Perlin marble, perlin/mod.rs:162-184: only the z lane survives the dot with (0,0,1);
        // the turbulence takes the UNSCALED point (:171).
        float turb = 10.0f * perlin_turbulence(&perlin[T.image], h.p);
        float noise = 0.5f * (1.0f + sinf(T.scalar * h.p.z + turb));
        return f3(noise, noise, noise);
    }
}

// ------------------------------------------------------------------------------------------
// Samplers (core/math.rs:33-81) — rejection loops in the reference's draw order
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 random_in_unit_sphere(Rng& rng) {   // math.rs:33-45
    for (;;) {
        float x = rng.gen_pm1(), y = rng.gen_pm1(), z = rng.gen_pm1();
        if (x * x + y * y + z * z <= 1.0f) return f3(x, y, z);
    }
}
__device__ __forceinline__ float3 unit(float3 v) {   // vec3.rs:170 (nalgebra normalize: divide by the norm)
    float inv = 1.0f / sqrtf(dot(v, v));
    return v * inv;
}

// Skybox::background (skybox/mod.rs:5-25)
__device__ __forceinline__ float3 background(const DeviceScene& s, float3 d) {
    if (s.sky_kind == B200RT_SKY_ABOVE) {
        float t = 0.5f * (d.y * rsqrtf(dot(d, d)) + 1.0f);
        return f3(1.0f - t + 0.5f * t, 1.0f - t + 0.7f * t, 1.0f);
    }
    return f3(s.sky_r, s.sky_g, s.sky_b);   // Flat(c); None is uploaded as Flat(0)
}

// Material::scatter + ::emitted (material_type.rs:50-79).  Returns false when the path
// ends (DiffuseLight).  `atten`/`emit` are the ray_color accumulators (render.rs:24-25).
struct ShadeOut { float3 o, d; bool scattered; };

template <class Acc>
__device__ __forceinline__ ShadeOut shade(const DeviceScene& s, const Acc& acc, const RayF& r, const HitRec& h,
                                          Rng& rng, float3& atten, float3& emit) {
    ShadeOut out;
    out.o = h.p;
    out.scattered = true;
    MatRec m = acc.mat(h.id);
    if (m.kind == B200RT_MAT_DIELECTRIC) {                              // dielectric.rs:22-49
        float ir = m.m0.w;
        float ratio = h.front ? 1.0f / ir : ir;
        float3 ud = unit(r.d);
        float cos_theta = fminf(-dot(ud, h.n), 1.0f);
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        bool reflect = ratio * sin_theta > 1.0f;
        if (!reflect) {                                                 // `||` short-circuit: no draw under TIR
            float r0 = (1.0f - ratio) / (1.0f + ratio);
            r0 = r0 * r0;
            float k = 1.0f - cos_theta;
            float k2 = k * k;
            float refl = r0 + (1.0f - r0) * (k2 * k2 * k);              // powf(5.0), dielectric.rs:18
            reflect = refl > rng.gen();
        }
        if (reflect) {
            out.d = ud - h.n * (2.0f * dot(ud, h.n));                   // vec3.rs:135
        } else {                                                        // vec3.rs:139-145
            float3 perp = (h.n * cos_theta + ud) * ratio;
            float par = -sqrtf(fabsf(1.0f - dot(perp, perp)));
            out.d = perp + h.n * par;
        }
        return out;                                                     // attenuation = ones
    }
    if (m.kind == B200RT_MAT_DIFFUSE_LIGHT) {                           // lighting.rs:21-28
        float3 e = m.tex >= 0 ? texture_value(acc, s.images, s.perlin, m.tex, h) : f3(m.m0.x, m.m0.y, m.m0.z);
        emit = emit + atten * e;
        out.scattered = false;
        out.d = r.d;
        return out;
    }
    // Metal, Lambertian and FairyLight all start from a point in the unit ball
    float3 rs = random_in_unit_sphere(rng);
    if (m.kind == B200RT_MAT_METAL) {                                   // metal.rs:27-39: never absorbs, sampler drawn even for fuzz 0
        float3 ud = unit(r.d);
        float3 refl = ud - h.n * (2.0f * dot(ud, h.n));
        out.d = refl + rs * m.m0.w;
        atten = atten * f3(m.m0.x, m.m0.y, m.m0.z);
        return out;
    }
    float3 a = m.tex >= 0 ? texture_value(acc, s.images, s.perlin, m.tex, h) : f3(m.m0.x, m.m0.y, m.m0.z);
    if (m.kind == B200RT_MAT_FAIRY_LIGHT) {                             // lighting.rs:59-66 then :43-57
        float scale = -dot(h.n, r.d) * rsqrtf(r.a);
        emit = emit + atten * (a * scale);
        a = unit(a);
    }
    float3 dir = h.n + unit(rs);                                        // lambertian.rs:23, math.rs:62-68
    if (fabsf(dir.x) < 1e-8f && fabsf(dir.y) < 1e-8f && fabsf(dir.z) < 1e-8f) dir = h.n;   // vec3.rs:130
    out.d = dir;
    atten = atten * a;
    return out;
}

// Camera::pixel_ray (camera/mod.rs:98-131); x, y are the jittered pixel coordinates.
__device__ __forceinline__ void pixel_ray(const DeviceCamera& cam, Rng& rng, float x, float y, float3* o, float3* d) {
    float3 offset = f3(0.0f, 0.0f, 0.0f);
    if (cam.has_lens) {                                                 // math.rs:70-81 random_in_unit_disk
        float rx, ry;
        do { rx = rng.gen_pm1(); ry = rng.gen_pm1(); } while (!(rx * rx + ry * ry <= 1.0f));
        offset = cam.ul * rx + cam.vl * ry;
    }
    *d = fma3(y, cam.vh, fma3(x, cam.hw, cam.llo)) - offset;
    *o = cam.origin + offset;
}

}  // namespace b200rt
