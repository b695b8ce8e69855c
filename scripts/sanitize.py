"""A small tour of every kernel instantiation in libb200rt.so (both accessors, chunked and whole-tile items, accumulate, rows /
shards, exact slabs, counters, both tree builders, the parity hooks), sized for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize.py
(compute-sanitizer is closed on this round's GPU pool, so it was only run plain there: 0.3 s, all asserts hold.)
Results are only checked for being finite / self-consistent — parity is tests/'s job."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shirley_raytracing_rs_b200 as rt

def rays_for(cam, n, seed):
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(0, cam.image_width, n), rng.uniform(0, cam.image_height, n)], axis=1).astype(np.float32)
    return rt.camera_rays(cam, xy, seed=seed)                                  # camera_rays_kernel

cam = rt.default_camera(96)
weekend = rt.Scene.named("random", seed=0xDEADBEEF)                            # staged in shared memory (SmemAcc)
r = rays_for(cam, 4096, 1)
ids, hits, _ = rt.closest_hit(weekend, r, 0.001, float("inf"))                 # closest_hit_kernel (K1)
assert (ids >= -1).all() and np.isfinite(hits["t"][ids >= 0]).all()
sc = rt.scatter(weekend, r[ids >= 0][:512], hits[ids >= 0][:512], seed=3)      # scatter_kernel (cooperative Perlin inside)
a1, st1 = rt.render(weekend, cam, samples=6, seed=1)                           # K2, tiles split into sample ranges + finalize_kernel
os.environ["B200RT_CHUNKS"] = "1"
a2, _ = rt.render(weekend, cam, samples=6, seed=1)                             # K2, whole tiles (direct write-back)
del os.environ["B200RT_CHUNKS"]
assert np.array_equal(a1, a2) and np.isfinite(a1).all()
a3, _ = rt.render(weekend, cam, samples=3, seed=1)
a3, _ = rt.render(weekend, cam, samples=3, seed=1, sample_offset=3, into=a3)   # accumulate flag
a4, _ = rt.render(weekend, rt.default_camera(50), samples=4, seed=2, shard=(3, 1), rows=(5, 29))   # ragged tiles, rows, shards
rgb = rt.resolve_rgb8(a1, samples=6)                                           # resolve_kernel
for name, param in (("cornell", 0), ("earth", 0), ("perlin", 0), ("random-night", 0), ("box-light", 0)):
    s = rt.Scene.named(name, seed=7, param=param)
    a, _ = rt.render(s, cam, samples=3, seed=4)
    assert np.isfinite(a).all(), name
big = rt.Scene.named("scaled", seed=3, param=40)                               # 6 400 spheres: read through L1 (GmemAcc)
ids_b, _, _ = rt.closest_hit(big, rays_for(cam, 2048, 5), 0.001, float("inf"))
ab, _ = rt.render(big, cam, samples=3, seed=6)
os.environ["B200RT_BUILDER"] = "lbvh"                                          # K5: the device-built linear BVH (CUB sort inside)
big2 = rt.Scene.named("scaled", seed=3, param=40)
ids_l, _, _ = rt.closest_hit(big2, rays_for(cam, 2048, 5), 0.001, float("inf"))
del os.environ["B200RT_BUILDER"]
assert np.array_equal(ids_b, ids_l)
os.environ["B200RT_FAST_SLAB"] = "0"                                           # exact Aabb::hit2 instantiation
ae, _ = rt.render(weekend, cam, samples=2, seed=1)
del os.environ["B200RT_FAST_SLAB"]
_, stc = rt.render(weekend, cam, samples=2, seed=1, count_traversal=True)      # COUNT instantiation
assert stc.node_visits > 0
boxes = np.array([[0, 0, 0, 1, 1, 1]], dtype=np.float32).repeat(64, 0)
rt.aabb_hit(boxes, r[:64])                                                     # aabb_hit_kernel
uvp = np.random.default_rng(9).uniform(-3, 3, (256, 5)).astype(np.float32); uvp[:, :2] = np.abs(uvp[:, :2]) / 3
for t in range(int(weekend.desc.contents.n_textures)):
    assert np.isfinite(rt.texture_value(weekend, t, uvp)).all()   # texture_value_kernel
rt.rng_uniforms(1, 2, 3, 64)                                                   # rng_kernel
print("sanitize tour ok:", int((ids >= 0).sum()), "hits,", st1.rays, "rays rendered")
