/* oracle.h — C ABI of the CPU oracle (TEST INFRASTRUCTURE ONLY; see oracle.hpp).
 * Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
 * The scene crosses as the same B200rtSceneDesc the product consumes (f32 scene data,
 * widened to f64 inside), so both sides see identical inputs. */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#include "../include/b200rt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleScene OracleScene;

typedef struct OracleHit {      /* geometry/hittable.rs:7-14 in f64 */
    double t, p[3], n[3], u, v;
    int32_t front_face, id;
} OracleHit;

/* How numerically decidable the closest hit of a ray is (all computed in f64 by testing
 * every object).  Tests drop rays whose margins are below a stated threshold before
 * demanding bit-exact ids between f32 GPU and f64 reference. */
typedef struct OracleMargin {
    double second_rel;   /* (t_second - t_best) / t_best over the other objects; +inf if none   */
    double graze;        /* min |discriminant| / (a r^2) over spheres in front of the origin    */
    double edge;         /* min relative distance of an in-range plane hit to a rect/box edge   */
    double tmin_rel;     /* min |root - t_min| / t_min over all roots of all objects            */
} OracleMargin;

typedef struct OracleScatter {
    double o[3], d[3], attenuation[3], emitted[3];
    int32_t scattered; uint32_t draws;
} OracleScatter;

typedef struct OracleStats {
    uint64_t rays, paths, depth_exhausted;
    uint64_t pops, box_hits, leaf_tests, max_stack;   /* bbox_tree.rs:68-88 counters */
    double seconds;
    int32_t threads;
    int32_t _pad;
} OracleStats;

/* reference_topology != 0: bvh/bbox_tree/constructor.rs:9-212 verbatim (O(N^2));
 * 0: median split (for scenes beyond a few thousand objects).  precision: 64 or 32. */
OracleScene* oracle_scene_create(const B200rtSceneDesc* desc, int reference_topology, int precision);
void oracle_scene_destroy(OracleScene* s);
int  oracle_tree_info(const OracleScene* s, uint64_t* n_nodes, uint64_t* max_depth);

/* Scene::hit via BboxTree::hit_workspace (scene/mod.rs:153-163, bbox_tree.rs:56-91). */
int  oracle_closest_hit(const OracleScene* s, const B200rtRay* rays, size_t n, double t_min, double t_max,
                        int32_t* ids, OracleHit* hits, OracleMargin* margins, OracleStats* stats);
/* f32 restatement of the device arithmetic over every object in id order (gpu_f32.hpp). */
int  oracle_closest_hit_f64(const OracleScene* s, const double* rays6, size_t n, double t_min, double t_max, int32_t* ids, OracleHit* hits);
int  oracle_closest_hit_gpu32(const B200rtSceneDesc* desc, const B200rtRay* rays, size_t n, float t_min, float t_max, B200rtHit* hits);

/* MaterialType::scatter + emitted. Record i draws from stream (seed, i, 0) unless
 * `injected` != NULL, in which case it consumes injected[i*stride .. ] as its uniforms. */
int  oracle_scatter(const OracleScene* s, const B200rtRay* rays, const OracleHit* hits, size_t n, uint64_t seed,
                    const double* injected, size_t stride, OracleScatter* out);
/* Camera::pixel_ray; record i draws the lens sample from stream (seed, i, 0). out = 6 doubles/ray. */
int  oracle_camera_rays(const B200rtCamera* cam, int precision, const double* xy, size_t n, uint64_t seed, double* out6);
int  oracle_texture_value(const OracleScene* s, int32_t tex, const double* uvp5, size_t n, double* out_rgb);
/* render_scanline over rows (OpenMP dynamic,1 over scanlines = rayon's row tasks,
 * src/main.rs:118-125).  accum = H*W*3 doubles, row 0 = bottom, SUM over samples.
 * threads <= 0: all cores.  Only rows [row_begin,row_end) (0,0 = all) are rendered. */
int  oracle_render(const OracleScene* s, const B200rtCamera* cam, const B200rtRenderParams* params, double* accum, OracleStats* stats, int threads);
/* to_image (image.rs:31-44) */
int  oracle_resolve_rgb8(const double* accum, uint32_t width, uint32_t height, uint32_t samples, uint8_t* out_rgb8);
int  oracle_rng_uniforms(uint64_t seed, uint32_t a, uint32_t b, size_t n, double* out);

/* KAT helpers for bvh/aabb.rs and core/fp.rs */
int  oracle_aabb_hit2(const double* box6, const double* ray6, double t_min, double t_max);
int  oracle_aabb_hit(const double* box6, const double* ray6, double t_min, double t_max);
void oracle_surrounding_box(const double* a6, const double* b6, double* out6);
double oracle_fmin(double a, double b);
double oracle_fmax(double a, double b);
uint64_t oracle_sizeof_tree_node_f64(void);

#ifdef __cplusplus
}
#endif
#endif
