"""K1 alone — the BVH-only workload of the reference's criterion bench (benches/my_benchmark.rs:35-75): a (2s)^3 lattice of
jittered log-normal spheres, random origins in the cube, uniform directions; ids only (no traversal counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shirley_raytracing_rs_b200 as rt

for side in (8, 32, 64):
    s = rt.Scene.named("lattice", seed=0xDEADBEEF, param=side)
    n = 4_000_000
    g = np.random.default_rng(1)
    d = g.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([g.uniform(-side, side, size=(n, 3)), d], axis=1).astype(np.float32)
    rt.closest_hit(s, rays[:1000])
    best = 1e9
    for rep in range(3):
        ids, _, st = rt.closest_hit(s, rays, want_hits=False)
        best = min(best, st.kernel_ms)
    _, _, sc = rt.closest_hit(s, rays[:200_000])
    print(f"K1 lattice side {side}: {s.desc.contents.n_prims} spheres, {n} rays: kernel {best:.2f} ms  {n / best / 1e3:.0f} Mrays/s  "
          f"hit fraction {(ids >= 0).mean():.3f}  node visits/ray {sc.node_visits / sc.rays:.1f}  prim tests/ray {sc.prim_tests / sc.rays:.1f}", flush=True)
