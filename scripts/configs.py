"""Run the BASELINE.json configs that are not the bench headline (C3 earth, C4 scaled scenes)
through the public API and print throughput: python scripts/configs.py [spp_scale]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shirley_raytracing_rs_b200 as rt

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(out_dir, exist_ok=True)


def run(name, scene, cam, spp, png=None):
    t0 = time.perf_counter(); scene.device(); t_create = time.perf_counter() - t0
    info = scene.info()
    rt.render(scene, cam, samples=2, seed=1)
    acc, st = rt.render(scene, cam, samples=spp, seed=2)
    print(f"{name}: {cam.image_width}x{cam.image_height} {spp} spp  objects {info.n_prims} bvh nodes {info.n_bvh_nodes} depth {info.bvh_depth} "
          f"nodes in smem {info.bvh_nodes_in_smem} device MB {info.device_bytes/1e6:.1f}  scene_create {t_create*1e3:.0f} ms | "
          f"kernel {st.kernel_ms:.1f} ms  {st.rays/st.kernel_ms/1e3:.0f} Mrays/s  segs/sample {st.rays/st.paths:.2f}", flush=True)
    if png:
        rt.write_png(os.path.join(out_dir, png), rt.resolve_rgb8(acc, samples=spp))


# C3: textured earth sphere + checker ground, 1920x1080, 256 spp (src/scenes.rs:81-93)
run("C3 earth", rt.Scene.named("earth"), rt.default_camera(1920, aspect_ratio=(16, 9)), max(1, int(256 * scale)), "c3_earth.png")
# C2-shaped frames of the other factories
run("perlin", rt.Scene.named("perlin"), rt.default_camera(1200), max(1, int(100 * scale)), "perlin.png")
run("cornell", rt.Scene.named("cornell"), rt.camera((278, 278, -800), (278, 278, 0), vfov=40, aperture=0.00001, width=800, aspect_ratio=(1, 1), focus_length=10.0), max(1, int(200 * scale)), "cornell.png")
run("random-night", rt.Scene.named("random-night", seed=7), rt.default_camera(1200), max(1, int(100 * scale)), "night.png")
# C4: scaled random-sphere scenes, 3840x2160, 64 spp
for G in (158, 500):
    t0 = time.perf_counter(); s = rt.Scene.named("scaled", seed=3, param=G); th = time.perf_counter() - t0
    cam = rt.camera((0.9 * G, 0.18 * G + 2, 0.35 * G), (0, 0, 0), vfov=30, aperture=0.001, width=3840, aspect_ratio=(16, 9), focus_length=10.0)
    print(f"  (host scene generation + flatten {th:.2f} s)")
    run(f"C4 scaled G={G}", s, cam, max(1, int(64 * scale)), f"c4_scaled_{G}.png")

# K1 alone (the criterion-shaped BVH microbenchmark): scripts/k1_microbench.py
