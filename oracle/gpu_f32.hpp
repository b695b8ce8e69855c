// gpu_f32.hpp — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// CPU restatement, in IEEE f32, of the arithmetic the device uses for primitive
// intersection (shirley_raytracing_rs_b200/csrc/rt_device.cuh: make_ray, sphere_roots,
// hit_sphere, rect_t, hit_rect, hit_box, accept_t, make_hit).  It is NOT a copy of the
// traversal: it tests every primitive in hit-id order (the reference's HitList::hit,
// scene/mod.rs:66-76, with the closest-so-far interval), which is what a BVH closest-hit
// must be equivalent to.  Same operations in the same order with explicit fmaf, so for a
// given ray the winning id, t, point and normal must match the GPU bit for bit
// (SURVEY.md §8c check 1: "GPU vs oracle<float> bit-exact on the full fixed ray set").
// The algorithm being restated is the reference's Sphere::hit / Rect::hit / RectBox::hit
// (geometry/sphere.rs:29-52, geometry/rect.rs:55-80,147-156) with the quadratic solved in
// its cancellation-free form and the box taken through its three slabs.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
#include <limits>

#include "../include/b200rt.h"

namespace oracle {
namespace gpu32 {

struct Ray32 { float ox, oy, oz, dx, dy, dz, ix, iy, iz, a, inv_a; };
inline Ray32 make_ray(const B200rtRay& r) {
    Ray32 q;
    q.ox = r.ox; q.oy = r.oy; q.oz = r.oz; q.dx = r.dx; q.dy = r.dy; q.dz = r.dz;
    q.ix = 1.0f / r.dx; q.iy = 1.0f / r.dy; q.iz = 1.0f / r.dz;
    q.a = fmaf(r.dz, r.dz, fmaf(r.dy, r.dy, r.dx * r.dx));
    q.inv_a = 1.0f / q.a;
    return q;
}
inline float comp_o(const Ray32& r, int i) { return i == 0 ? r.ox : (i == 1 ? r.oy : r.oz); }
inline float comp_d(const Ray32& r, int i) { return i == 0 ? r.dx : (i == 1 ? r.dy : r.dz); }
inline float comp_i(const Ray32& r, int i) { return i == 0 ? r.ix : (i == 1 ? r.iy : r.iz); }

struct Closest32 { float t; int id; uint32_t type; int face; };

inline bool accept_t(float t, float t_min, const Closest32& c, int id) {
    return (t >= t_min) && (t < c.t || (t == c.t && (c.id < 0 || id > c.id)));
}

inline void hit_sphere(const Ray32& r, const B200rtSphere& s, int id, float t_min, Closest32& c) {
    float ocx = r.ox - s.cx, ocy = r.oy - s.cy, ocz = r.oz - s.cz;
    float bp = -fmaf(ocz, r.dz, fmaf(ocy, r.dy, ocx * r.dx));
    float q = bp * r.inv_a;
    float lx = fmaf(q, r.dx, ocx), ly = fmaf(q, r.dy, ocy), lz = fmaf(q, r.dz, ocz);
    float l2 = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
    float r2 = s.radius * s.radius;
    float delta = r2 - l2;
    if (delta < 0.0f) return;
    float sq = sqrtf(delta * r.a);
    float qq = bp + copysignf(sq, bp);
    float cc = fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx)) - r2;
    float r0 = cc / qq;
    float r1 = qq * r.inv_a;
    float lo = fminf(r0, r1), hi = fmaxf(r0, r1);
    float root = lo;
    bool ok = accept_t(root, t_min, c, id);
    if (!ok) { root = hi; ok = accept_t(root, t_min, c, id); }
    if (ok) { c.t = root; c.id = id; c.type = B200RT_PRIM_SPHERE; }
}

inline float rect_t(const Ray32& r, int d1, int d2, float d1_min, float d1_max, float d2_min, float d2_max, float offset, float t_min, float t_max) {
    int dn = 3 - d1 - d2;
    float t = (offset - comp_o(r, dn)) * comp_i(r, dn);
    if (!(t >= t_min && t <= t_max)) return std::numeric_limits<float>::quiet_NaN();
    float a = fmaf(t, comp_d(r, d1), comp_o(r, d1));
    float b = fmaf(t, comp_d(r, d2), comp_o(r, d2));
    if (a < d1_min || a > d1_max || b < d2_min || b > d2_max) return std::numeric_limits<float>::quiet_NaN();
    return t;
}
inline void rect_axes(uint32_t type, int& d1, int& d2) {
    d1 = (type == B200RT_PRIM_RECT_YZ) ? 1 : 0;
    d2 = (type == B200RT_PRIM_RECT_XY) ? 1 : 2;
}
inline void hit_rect(const Ray32& r, const B200rtRect& g, uint32_t type, int id, float t_min, Closest32& c) {
    int d1, d2; rect_axes(type, d1, d2);
    float t = rect_t(r, d1, d2, g.d1_min, g.d1_max, g.d2_min, g.d2_max, g.offset, t_min, c.t);
    if (accept_t(t, t_min, c, id)) { c.t = t; c.id = id; c.type = type; }
}
inline void hit_box(const Ray32& r, const B200rtBox& g, int id, float t_min, Closest32& c) {
    // RectBox::hit (rect.rs:147-156) through its three slabs: entry point if inside
    // [t_min, closest], else exit point; axis priority y > x > z at equal t ("later face
    // replaces", rect.rs:149-154).
    const float* lo = g.min; const float* hi = g.max;
    float ax = (lo[0] - r.ox) * r.ix, bx = (hi[0] - r.ox) * r.ix;
    float ay = (lo[1] - r.oy) * r.iy, by = (hi[1] - r.oy) * r.iy;
    float az = (lo[2] - r.oz) * r.iz, bz = (hi[2] - r.oz) * r.iz;
    float nx = fminf(ax, bx), fx = fmaxf(ax, bx);
    float ny = fminf(ay, by), fy = fmaxf(ay, by);
    float nz = fminf(az, bz), fz = fmaxf(az, bz);
    float t_enter = fmaxf(fmaxf(nx, ny), nz), t_exit = fminf(fminf(fx, fy), fz);
    if (!(t_enter <= t_exit)) return;
    bool entering = t_enter >= t_min;
    float t = entering ? t_enter : t_exit;
    if (!accept_t(t, t_min, c, id)) return;
    float ty = entering ? ny : fy, tx = entering ? nx : fx;
    c.t = t; c.id = id; c.type = B200RT_PRIM_BOX;
    c.face = (ty == t) ? 4 : ((tx == t) ? 2 : 0);
}

// Linear closest hit over every object in id order; fills `out` like the device make_hit.
inline void closest_hit(const B200rtSceneDesc& d, const B200rtRay& ray, float t_min, float t_max, B200rtHit* out) {
    Ray32 r = make_ray(ray);
    Closest32 c; c.t = t_max; c.id = -1; c.type = 0; c.face = 0;
    for (uint32_t i = 0; i < d.n_prims; ++i) {
        const B200rtPrimRef& p = d.prims[i];
        if (p.type == B200RT_PRIM_SPHERE) {
            const B200rtSphere& s = d.spheres[p.index];
            // a negative radius has an inverted bounding box that Aabb::hit2 never passes
            // (sphere.rs:54-60, aabb.rs:62-79): the object cannot be hit through the tree
            if (!(s.cx - s.radius <= s.cx + s.radius)) continue;
            hit_sphere(r, s, (int)i, t_min, c);
        } else if (p.type == B200RT_PRIM_BOX) {
            const B200rtBox& b = d.boxes[p.index];
            if (!(b.min[0] <= b.max[0] && b.min[1] <= b.max[1] && b.min[2] <= b.max[2])) continue;
            hit_box(r, b, (int)i, t_min, c);
        } else {
            const B200rtRect& g = d.rects[p.index];
            if (!(g.d1_min <= g.d1_max && g.d2_min <= g.d2_max)) continue;
            hit_rect(r, g, p.type, (int)i, t_min, c);
        }
    }
    B200rtHit h; std::memset(&h, 0, sizeof h);
    h.id = c.id;
    if (c.id >= 0) {
        h.t = c.t;
        h.p[0] = fmaf(c.t, r.dx, r.ox); h.p[1] = fmaf(c.t, r.dy, r.oy); h.p[2] = fmaf(c.t, r.dz, r.oz);
        float n[3];
        if (c.type == B200RT_PRIM_SPHERE) {
            const B200rtSphere& s = d.spheres[d.prims[c.id].index];
            float inv_r = 1.0f / s.radius;
            float ocx = r.ox - s.cx, ocy = r.oy - s.cy, ocz = r.oz - s.cz;
            n[0] = fmaf(c.t, r.dx, ocx) * inv_r; n[1] = fmaf(c.t, r.dy, ocy) * inv_r; n[2] = fmaf(c.t, r.dz, ocz) * inv_r;
        } else {
            int dn;
            if (c.type == B200RT_PRIM_BOX) dn = c.face < 2 ? 2 : (c.face < 4 ? 0 : 1);
            else { int d1, d2; rect_axes(c.type, d1, d2); dn = 3 - d1 - d2; }
            n[0] = dn == 0 ? 1.0f : 0.0f; n[1] = dn == 1 ? 1.0f : 0.0f; n[2] = dn == 2 ? 1.0f : 0.0f;
        }
        float dnv = fmaf(r.dz, n[2], fmaf(r.dy, n[1], r.dx * n[0]));
        bool front = dnv < 0.0f;
        h.front_face = front ? 1 : 0;
        for (int k = 0; k < 3; ++k) h.n[k] = front ? n[k] : -n[k];
    }
    *out = h;
}

}  // namespace gpu32
}  // namespace oracle
