"""BASELINE config 4 (scaled random-sphere scene, G=500 -> ~1e6 spheres, 3840x2160): one frame at a few spp,
for ncu (`-k regex:path_trace`).  Prints the traversal counters used for the memory roofline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shirley_raytracing_rs_b200 as rt
G = int(sys.argv[1]) if len(sys.argv) > 1 else 500
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = rt.Scene.named("scaled", seed=3, param=G)
cam = rt.camera((0.9 * G, 0.18 * G + 2, 0.35 * G), (0, 0, 0), vfov=30, aperture=0.001, width=3840, aspect_ratio=(16, 9), focus_length=10.0)
rt.render(s, cam, samples=1, seed=1)
_, st = rt.render(s, cam, samples=spp, seed=2)
_, sc = rt.render(s, cam, samples=spp, seed=2, count_traversal=True)
info = s.info()
nv, pt = sc.node_visits / sc.rays, sc.prim_tests / sc.rays
bytes_per_ray = 64 * nv + 16 * pt + 32
print(f"C4 G={G}: objects {info.n_prims} nodes {info.n_bvh_nodes} depth {info.bvh_depth} top nodes in smem {info.bvh_nodes_in_smem} device MB {info.device_bytes / 1e6:.1f}  "
      f"builder {'device LBVH' if info.bvh_builder else 'host SAH'} {info.bvh_build_ms:.1f} ms")
print(f"  {spp} spp: kernel {st.kernel_ms:.1f} ms  {st.rays / st.kernel_ms / 1e3:.0f} Mrays/s  node visits/ray {nv:.2f}  prim tests/ray {pt:.2f}  "
      f"algorithmic bytes/ray {bytes_per_ray:.0f} -> {st.rays / st.kernel_ms / 1e6 * bytes_per_ray:.0f} GB/s")
