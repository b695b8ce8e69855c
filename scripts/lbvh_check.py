import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import shirley_raytracing_rs_b200 as rt
from oracle import pyoracle as po
from common import random_rays
# parity: ids bit-exact vs the f32 mirror with the device-built tree
for name, param, scale in (("scaled", 40, 45.0), ("lattice", 4, 6.0), ("random", 0, 8.0)):
    os.environ["B200RT_BUILDER"] = "lbvh"
    s = rt.Scene.named(name, seed=11, param=param)
    rays = random_rays(200_000, 5, origin_scale=scale)
    ids, hits, st = rt.closest_hit(s, rays, 0.001, float("inf"))
    want = po.closest_hit_gpu32(s.desc, rays, 0.001, float("inf"))
    info = s.info()
    hit = ids >= 0
    print(name, "prims", info.n_prims, "depth", info.bvh_depth, "builder", info.bvh_builder, f"{info.bvh_build_ms:.2f} ms", "ids equal", np.array_equal(ids, want["id"]), "t equal", np.array_equal(hits["t"][hit], want["t"][hit]), "hit frac", hit.mean(), "nodes/ray", st.node_visits / st.rays)
    os.environ["B200RT_BUILDER"] = "sah"
    s2 = rt.Scene.named(name, seed=11, param=param)
    ids2, _, st2 = rt.closest_hit(s2, rays, 0.001, float("inf"))
    print("   sah: depth", s2.info().bvh_depth, f"{s2.info().bvh_build_ms:.2f} ms", "ids equal", np.array_equal(ids2, ids), "nodes/ray", st2.node_visits / st2.rays)
