// oracle_capi.cpp — C ABI over oracle.hpp / gpu_f32.hpp (TEST INFRASTRUCTURE ONLY).
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <memory>

#include "gpu_f32.hpp"
#include "oracle.h"
#include "oracle.hpp"

using namespace oracle;

struct OracleScene {
    int precision = 64;
    std::unique_ptr<Scene<double>> s64;
    std::unique_ptr<Scene<float>> s32;
};

namespace {

template <class Real> Ray<Real> to_ray(const B200rtRay& r) {
    return {{(Real)r.ox, (Real)r.oy, (Real)r.oz}, {(Real)r.dx, (Real)r.dy, (Real)r.dz}};
}
template <class Real> void put_hit(const HitRecord<Real>& h, long id, OracleHit* o) {
    o->t = (double)h.t;
    o->p[0] = (double)h.point.x; o->p[1] = (double)h.point.y; o->p[2] = (double)h.point.z;
    o->n[0] = (double)h.normal.x; o->n[1] = (double)h.normal.y; o->n[2] = (double)h.normal.z;
    o->u = (double)h.u; o->v = (double)h.v; o->front_face = h.front_face ? 1 : 0; o->id = (int32_t)id;
}

// Decidability margins in f64 by testing every object (no BVH).
void margins_for(const Scene<double>& sc, const Ray<double>& ray, double t_min, double t_max, long best_id, double best_t, OracleMargin* m) {
    const double INF = std::numeric_limits<double>::infinity();
    double second = INF, graze = INF, edge = INF, tmin_rel = INF;
    auto rect_edge = [&](const Rect<double>& r) {
        int dn = 3 - r.d1 - r.d2;
        double t = (r.offset - ray.orig[dn]) / ray.direction[dn];
        if (!(t >= t_min * 0.5 && t <= t_max)) return;
        double a = ray.orig[r.d1] + t * ray.direction[r.d1], b = ray.orig[r.d2] + t * ray.direction[r.d2];
        double e1 = std::min(std::fabs(a - r.d1_min), std::fabs(a - r.d1_max)) / std::max(1.0, std::fabs(r.d1_max - r.d1_min));
        double e2 = std::min(std::fabs(b - r.d2_min), std::fabs(b - r.d2_max)) / std::max(1.0, std::fabs(r.d2_max - r.d2_min));
        // only an edge that could flip inside/outside matters: the other coordinate must be (nearly) inside
        bool in1 = a >= r.d1_min - 1e-3 && a <= r.d1_max + 1e-3, in2 = b >= r.d2_min - 1e-3 && b <= r.d2_max + 1e-3;
        if (in2) edge = std::min(edge, e1);
        if (in1) edge = std::min(edge, e2);
        tmin_rel = std::min(tmin_rel, std::fabs(t - t_min) / t_min);
    };
    for (size_t i = 0; i < sc.objects.size(); ++i) {
        const GeometricObject<double>& g = sc.objects[i];
        if (g.type == B200RT_PRIM_SPHERE) {
            const Sphere<double>& s = g.sphere;
            if (!(s.radius > 0)) continue;
            Vec3<double> oc = ray.orig - s.center;
            double a = ray.direction.length_squared();
            double hb = oc.dot(ray.direction);
            double q = -hb / a;                                  // closest-approach parameter
            Vec3<double> l = oc + ray.direction.scale(q);
            double delta_n = (s.radius * s.radius - l.length_squared()) / (s.radius * s.radius);
            double sq = delta_n >= 0 ? std::sqrt(delta_n * s.radius * s.radius / a) : 0.0;
            double far_root = q + sq;
            if (far_root > 0.0 && q - sq < t_max) graze = std::min(graze, std::fabs(delta_n));
            if (delta_n >= 0) {
                tmin_rel = std::min(tmin_rel, std::min(std::fabs(q - sq - t_min), std::fabs(q + sq - t_min)) / t_min);
            }
        } else if (g.type == B200RT_PRIM_BOX) {
            for (int k = 0; k < 6; ++k) rect_edge(g.box.sides[k]);
        } else {
            rect_edge(g.rect);
        }
        if ((long)i == best_id) continue;
        HitRecord<double> h;
        if (g.bounding_box().min.x <= g.bounding_box().max.x && g.hit(ray, t_min, t_max, h)) second = std::min(second, h.t);
    }
    m->second_rel = (best_id >= 0 && second < INF) ? (second - best_t) / best_t : INF;
    m->graze = graze; m->edge = edge; m->tmin_rel = tmin_rel;
}

template <class Real>
int closest_hit_impl(const Scene<Real>& sc, const Scene<double>* sc64, const B200rtRay* rays, size_t n, double t_min, double t_max,
                     int32_t* ids, OracleHit* hits, OracleMargin* margins, OracleStats* stats) {
    TraversalCounters total;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
    {
        std::vector<size_t> stack;
        TraversalCounters ctr;
#pragma omp for schedule(dynamic, 1024)
        for (long i = 0; i < (long)n; ++i) {
            Ray<Real> ray = to_ray<Real>(rays[i]);
            HitRecord<Real> rec;
            long id = sc.hit(stack, ray, (Real)t_min, (Real)t_max, rec, &ctr);
            ids[i] = (int32_t)id;
            if (hits) { if (id >= 0) put_hit(rec, id, &hits[i]); else { std::memset(&hits[i], 0, sizeof(OracleHit)); hits[i].id = -1; } }
            if (margins && sc64) margins_for(*sc64, to_ray<double>(rays[i]), t_min, t_max, id, id >= 0 ? (double)rec.t : 0.0, &margins[i]);
        }
#pragma omp critical
        { total.pops += ctr.pops; total.box_hits += ctr.box_hits; total.leaf_tests += ctr.leaf_tests; total.max_stack = std::max(total.max_stack, ctr.max_stack); }
    }
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->rays = n; stats->pops = total.pops; stats->box_hits = total.box_hits; stats->leaf_tests = total.leaf_tests; stats->max_stack = total.max_stack;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->threads = omp_get_max_threads();
    }
    return 0;
}

template <class Real>
int scatter_impl(const Scene<Real>& sc, const B200rtRay* rays, const OracleHit* hits, size_t n, uint64_t seed, const double* injected, size_t stride, OracleScatter* out) {
    for (size_t i = 0; i < n; ++i) {
        OracleScatter o; std::memset(&o, 0, sizeof o);
        const OracleHit& hi = hits[i];
        if (hi.id >= 0 && (size_t)hi.id < sc.objects.size()) {
            Ray<Real> ray = to_ray<Real>(rays[i]);
            HitRecord<Real> rec;
            rec.point = {(Real)hi.p[0], (Real)hi.p[1], (Real)hi.p[2]};
            rec.normal = {(Real)hi.n[0], (Real)hi.n[1], (Real)hi.n[2]};
            rec.t = (Real)hi.t; rec.u = (Real)hi.u; rec.v = (Real)hi.v; rec.front_face = hi.front_face != 0;
            UniformSource us;
            if (injected) { us.injected = injected + i * stride; us.n_injected = stride; } else us.rng = Rng(seed, (uint32_t)i, 0u);
            Vec3<Real> e;
            if (sc.emitted(hi.id, ray, rec, e)) { o.emitted[0] = (double)e.x; o.emitted[1] = (double)e.y; o.emitted[2] = (double)e.z; }
            Scatter<Real> s;
            if (sc.scatter(hi.id, us, ray, rec, s)) {
                o.scattered = 1;
                o.o[0] = (double)s.direction.orig.x; o.o[1] = (double)s.direction.orig.y; o.o[2] = (double)s.direction.orig.z;
                o.d[0] = (double)s.direction.direction.x; o.d[1] = (double)s.direction.direction.y; o.d[2] = (double)s.direction.direction.z;
                o.attenuation[0] = (double)s.attenuation.x; o.attenuation[1] = (double)s.attenuation.y; o.attenuation[2] = (double)s.attenuation.z;
            }
            o.draws = us.draws();
        }
        out[i] = o;
    }
    return 0;
}

template <class Real>
int render_impl(const Scene<Real>& sc, const B200rtCamera* cam_in, const B200rtRenderParams* prm, double* accum, OracleStats* stats, int threads) {
    Camera<Real> cam(*cam_in);
    uint32_t H = cam.H, W = cam.W;
    uint32_t r0 = prm->row_begin, r1 = prm->row_end;
    if (r0 == 0 && r1 == 0) r1 = H;
    if (r1 > H || r0 > r1) return -1;
    size_t samples = prm->samples == 0 ? 1 : prm->samples;   // src/main.rs:75-80
    if (threads <= 0) threads = omp_get_max_threads();
    RenderCounters total;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(threads)
    {
        std::vector<size_t> stack;           // BboxTreeWorkspace, one per worker (src/main.rs:119)
        std::vector<Vec3<Real>> line(W);
        RenderCounters ctr;
#pragma omp for schedule(dynamic, 1)
        for (long j = (long)r0; j < (long)r1; ++j) {
            render_scanline<Real>(cam, sc, prm->seed, prm->sample_offset, samples, prm->max_depth, stack, (size_t)j, line.data(), &ctr);
            for (uint32_t i = 0; i < W; ++i) {
                double* px = accum + ((size_t)j * W + i) * 3;
                px[0] = (double)line[i].x; px[1] = (double)line[i].y; px[2] = (double)line[i].z;
            }
        }
#pragma omp critical
        {
            total.rays += ctr.rays; total.paths += ctr.paths; total.depth_exhausted += ctr.depth_exhausted;
            total.trav.pops += ctr.trav.pops; total.trav.box_hits += ctr.trav.box_hits; total.trav.leaf_tests += ctr.trav.leaf_tests;
            total.trav.max_stack = std::max(total.trav.max_stack, ctr.trav.max_stack);
        }
    }
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->rays = total.rays; stats->paths = total.paths; stats->depth_exhausted = total.depth_exhausted;
        stats->pops = total.trav.pops; stats->box_hits = total.trav.box_hits; stats->leaf_tests = total.trav.leaf_tests; stats->max_stack = total.trav.max_stack;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->threads = threads;
    }
    return 0;
}

template <class Real> uint64_t tree_depth(const BboxTree<Real>& t) {
    if (!t.has_root) return 0;
    uint64_t best = 0;
    std::vector<std::pair<size_t, uint64_t>> st;
    st.push_back({t.root, 1});
    while (!st.empty()) {
        auto [n, d] = st.back(); st.pop_back();
        best = std::max(best, d);
        if (!t.tree[n].leaf) { st.push_back({t.tree[n].lhs, d + 1}); st.push_back({t.tree[n].rhs, d + 1}); }
    }
    return best;
}

}  // namespace

extern "C" {

OracleScene* oracle_scene_create(const B200rtSceneDesc* desc, int reference_topology, int precision) {
    if (!desc || (precision != 64 && precision != 32)) return nullptr;
    OracleScene* s = new OracleScene();
    s->precision = precision;
    // the f64 scene always exists: it also serves the margin computation
    s->s64.reset(new Scene<double>());
    s->s64->load(*desc, reference_topology != 0);
    if (precision == 32) { s->s32.reset(new Scene<float>()); s->s32->load(*desc, reference_topology != 0); }
    return s;
}
void oracle_scene_destroy(OracleScene* s) { delete s; }

int oracle_tree_info(const OracleScene* s, uint64_t* n_nodes, uint64_t* max_depth) {
    if (!s) return -1;
    if (n_nodes) *n_nodes = s->s64->tree.tree.size();
    if (max_depth) *max_depth = tree_depth(s->s64->tree);
    return 0;
}

int oracle_closest_hit(const OracleScene* s, const B200rtRay* rays, size_t n, double t_min, double t_max, int32_t* ids, OracleHit* hits, OracleMargin* margins, OracleStats* stats) {
    if (!s || (n && (!rays || !ids))) return -1;
    if (s->precision == 32) return closest_hit_impl<float>(*s->s32, s->s64.get(), rays, n, t_min, t_max, ids, hits, margins, stats);
    return closest_hit_impl<double>(*s->s64, s->s64.get(), rays, n, t_min, t_max, ids, hits, margins, stats);
}

// The same query with f64 rays (the reference's own precision, core/math.rs:5): what pins the restatement against
// the reference's unit-test values exactly — a ray such as (0.9, 0.9, -1.5) is not representable in f32.
int oracle_closest_hit_f64(const OracleScene* s, const double* rays6, size_t n, double t_min, double t_max, int32_t* ids, OracleHit* hits) {
    if (!s || !s->s64 || (n && (!rays6 || !ids))) return -1;
    std::vector<size_t> stack;
    for (size_t i = 0; i < n; ++i) {
        const double* r = rays6 + 6 * i;
        Ray<double> ray{{r[0], r[1], r[2]}, {r[3], r[4], r[5]}};
        HitRecord<double> rec;
        long id = s->s64->hit(stack, ray, t_min, t_max, rec, nullptr);
        ids[i] = (int32_t)id;
        if (hits) { if (id >= 0) put_hit(rec, id, &hits[i]); else { std::memset(&hits[i], 0, sizeof(OracleHit)); hits[i].id = -1; } }
    }
    return 0;
}

int oracle_closest_hit_gpu32(const B200rtSceneDesc* desc, const B200rtRay* rays, size_t n, float t_min, float t_max, B200rtHit* hits) {
    if (!desc || (n && (!rays || !hits))) return -1;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; ++i) gpu32::closest_hit(*desc, rays[i], t_min, t_max, &hits[i]);
    return 0;
}

int oracle_scatter(const OracleScene* s, const B200rtRay* rays, const OracleHit* hits, size_t n, uint64_t seed, const double* injected, size_t stride, OracleScatter* out) {
    if (!s || (n && (!rays || !hits || !out))) return -1;
    if (s->precision == 32) return scatter_impl<float>(*s->s32, rays, hits, n, seed, injected, stride, out);
    return scatter_impl<double>(*s->s64, rays, hits, n, seed, injected, stride, out);
}

int oracle_camera_rays(const B200rtCamera* cam, int precision, const double* xy, size_t n, uint64_t seed, double* out6) {
    if (!cam || (n && (!xy || !out6))) return -1;
    for (size_t i = 0; i < n; ++i) {
        UniformSource us; us.rng = Rng(seed, (uint32_t)i, 0u);
        if (precision == 32) {
            Camera<float> c(*cam);
            Ray<float> r = c.pixel_ray(us, (float)xy[2 * i], (float)xy[2 * i + 1]);
            double v[6] = {r.orig.x, r.orig.y, r.orig.z, r.direction.x, r.direction.y, r.direction.z};
            std::memcpy(out6 + 6 * i, v, sizeof v);
        } else {
            Camera<double> c(*cam);
            Ray<double> r = c.pixel_ray(us, xy[2 * i], xy[2 * i + 1]);
            double v[6] = {r.orig.x, r.orig.y, r.orig.z, r.direction.x, r.direction.y, r.direction.z};
            std::memcpy(out6 + 6 * i, v, sizeof v);
        }
    }
    return 0;
}

int oracle_texture_value(const OracleScene* s, int32_t tex, const double* uvp5, size_t n, double* out_rgb) {
    if (!s || (n && (!uvp5 || !out_rgb))) return -1;
    if (tex < 0 || (size_t)tex >= s->s64->textures.size()) return -1;
    for (size_t i = 0; i < n; ++i) {
        const double* q = uvp5 + 5 * i;
        if (s->precision == 32) {
            Vec3<float> c = s->s32->texture_value(tex, (float)q[0], (float)q[1], {(float)q[2], (float)q[3], (float)q[4]});
            out_rgb[3 * i] = c.x; out_rgb[3 * i + 1] = c.y; out_rgb[3 * i + 2] = c.z;
        } else {
            Vec3<double> c = s->s64->texture_value(tex, q[0], q[1], {q[2], q[3], q[4]});
            out_rgb[3 * i] = c.x; out_rgb[3 * i + 1] = c.y; out_rgb[3 * i + 2] = c.z;
        }
    }
    return 0;
}

int oracle_render(const OracleScene* s, const B200rtCamera* cam, const B200rtRenderParams* params, double* accum, OracleStats* stats, int threads) {
    if (!s || !cam || !params || !accum) return -1;
    if (s->precision == 32) return render_impl<float>(*s->s32, cam, params, accum, stats, threads);
    return render_impl<double>(*s->s64, cam, params, accum, stats, threads);
}

int oracle_resolve_rgb8(const double* accum, uint32_t width, uint32_t height, uint32_t samples, uint8_t* out_rgb8) {
    if (!accum || !out_rgb8 || samples == 0) return -1;
    std::vector<Vec3<double>> data((size_t)width * height);
    for (size_t i = 0; i < data.size(); ++i) data[i] = {accum[3 * i], accum[3 * i + 1], accum[3 * i + 2]};
    to_image<double>(data.data(), width, height, samples, out_rgb8);
    return 0;
}

int oracle_rng_uniforms(uint64_t seed, uint32_t a, uint32_t b, size_t n, double* out) {
    Rng r(seed, a, b);
    for (size_t i = 0; i < n; ++i) out[i] = r.gen<double>();
    return 0;
}

static Aabb<double> box_of(const double* b) { return {{b[0], b[1], b[2]}, {b[3], b[4], b[5]}}; }
static Ray<double> ray_of(const double* r) { return {{r[0], r[1], r[2]}, {r[3], r[4], r[5]}}; }
int oracle_aabb_hit2(const double* box6, const double* ray6, double t_min, double t_max) { return box_of(box6).hit2(ray_of(ray6), t_min, t_max) ? 1 : 0; }
int oracle_aabb_hit(const double* box6, const double* ray6, double t_min, double t_max) { return box_of(box6).hit(ray_of(ray6), t_min, t_max) ? 1 : 0; }
void oracle_surrounding_box(const double* a6, const double* b6, double* out6) {
    Aabb<double> r = surrounding_box(box_of(a6), box_of(b6));
    out6[0] = r.min.x; out6[1] = r.min.y; out6[2] = r.min.z; out6[3] = r.max.x; out6[4] = r.max.y; out6[5] = r.max.z;
}
double oracle_fmin(double a, double b) { return fmin_<double>(a, b); }
double oracle_fmax(double a, double b) { return fmax_<double>(a, b); }
uint64_t oracle_sizeof_tree_node_f64(void) { return sizeof(TreeNode<double>); }

}  // extern "C"
