"""Small invocations of every kernel for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shirley_raytracing_rs_b200 as rt
for name, param in (("random", 0), ("cornell", 0), ("earth", 0), ("scaled", 30)):
    s = rt.Scene.named(name, seed=5, param=param)
    cam = rt.default_camera(72)
    for env in ({"B200RT_KERNEL": "2"}, {"B200RT_KERNEL": "3"}, {"B200RT_KERNEL": "1"}, {"B200RT_KERNEL": "2", "B200RT_FAST_SLAB": "0"}):
        os.environ.update(env)
        acc, st = rt.render(s, cam, samples=3, seed=1, count_traversal=True)
        acc, st = rt.render(s, cam, samples=2, seed=1, rows=(3, 41), shard=(2, 1))
    rays = np.random.default_rng(0).normal(size=(5000, 6)).astype(np.float32)
    ids, hits, _ = rt.closest_hit(s, rays)
    sel = ids >= 0
    if sel.any():
        rt.scatter(s, rays[sel], hits[sel], seed=3)
    rt.texture_value(s, 0, np.random.default_rng(1).uniform(-3, 3, size=(1000, 5)).astype(np.float32)) if s.desc.contents.n_textures else None
    rgb = rt.resolve_rgb8(acc, samples=2)
rt.camera_rays(rt.default_camera(64), np.zeros((100, 2), np.float32))
rt.aabb_hit(np.zeros((10, 6), np.float32), np.ones((10, 6), np.float32))
rt.rng_uniforms(1, 2, 3, 16)
print("sanitize_small done")
