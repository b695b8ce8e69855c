#!/usr/bin/env python
"""Regenerate the golden fixtures of tests/golden/ (run here, on CPU: python tests/golden/make_golden.py).

The reference is a Rust crate that cannot be built in this image (no rustc/cargo), so these
vectors do not come from the reference binary.  They pin the two things the parity tests lean on:

  reference_kats.json      the known answers of the reference's OWN unit tests for this path
                           (bvh/bbox_tree.rs:122-227, bvh/aabb.rs:94-166, core/fp.rs:35-112), transcribed
  weekend_scene.json       the BASELINE config-1/2 scene in the reference's serde wire format
                           (host generator, seed 0xDEADBEEF) — so the vectors below do not depend on the
                           generator staying bit-stable
  weekend_hits.npz         a fixed ray set over that scene with the f64 ORACLE's closest-hit answers
                           (id, t, normal, decidability margins) and scatter outputs for the documented RNG stream
  weekend_mean_48x32.npy   the oracle's converged linear mean image (f64, 8192 spp) of a 48x32 frame

tests/test_golden.py checks the oracle against them on CPU (drift guard) and the CUDA path on the GPU."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import shirley_raytracing_rs_b200 as rt   # noqa: E402  (host side only: scene generator + camera)
from oracle import pyoracle as po          # noqa: E402
from common import fixed_ray_set           # noqa: E402

KATS = {
    "source": "reference unit tests: src/raytracer/bvh/bbox_tree.rs, bvh/aabb.rs, core/fp.rs",
    "closest_hit": [
        {"name": "KA1 miss_single_sphere (bbox_tree.rs:122-131)", "spheres": [[[0, 0, -10], 0.5]], "ray": [0, 0, 0, 1, 0, 0], "t_min": 0.0, "id": -1},
        {"name": "KA2 hit_single_sphere (bbox_tree.rs:134-148)", "spheres": [[[0, 0, -10], 0.5]], "ray": [0, 0, 0, 0, 0, -1], "t_min": 0.0,
         "id": 0, "t": 9.5, "p": [0, 0, -9.5], "n": [0, 0, 1], "front_face": True},
        {"name": "KA3 hit_box_but_miss_sphere (bbox_tree.rs:150-171)", "spheres": [[[0, 0, -2], 1.0]], "ray": [0, 0, 0, 0.9, 0.9, -1.5], "t_min": 0.0, "id": -1},
        {"name": "KA4 first of a chain of 100 (bbox_tree.rs:174-195)", "spheres": [[[0, 0, -2.0 * k], 1.0] for k in range(1, 101)], "ray": [0, 0, 0, 0, 0, -1],
         "t_min": 0.0, "id": 0, "t": 1.0},
        {"name": "KA5 object behind the first box (bbox_tree.rs:198-227)", "spheres": [[[0, 0, -2], 1.0], [[2, 2, -4], 1.0]], "ray": [0, 0, 0, 0.9, 0.9, -1.5], "t_min": 0.0,
         "id": 1, "t": 2.022009319139565, "n": [-0.18019161277439122, -0.18019161277439122, 0.9669860212906523]},
    ],
    "aabb_hit2": [
        {"name": "KA6 hit (aabb.rs:136-145)", "box": [1, -1, -1, 2, 1, 1], "ray": [0, 0, 0, 1, 0, 0], "t_min": 0.0, "t_max": 1e308, "hit": True},
        {"name": "KA6 miss (aabb.rs:147-156)", "box": [1, -1, -1, 2, 1, 1], "ray": [0, 2, 2, 1, 0, 0], "t_min": 0.0, "t_max": 1e308, "hit": False},
        {"name": "KA6 graze counts as a hit: 0 * inf = NaN leaves the interval alone (aabb.rs:158-166)", "box": [1, -1, -1, 2, 1, 1], "ray": [0, 1, 1, 1, 0, 0],
         "t_min": 0.0, "t_max": 1e308, "hit": True},
    ],
}


def main():
    json.dump(KATS, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)

    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    open(os.path.join(HERE, "weekend_scene.json"), "w").write(scene.to_json())
    scene = rt.Scene.from_json(open(os.path.join(HERE, "weekend_scene.json")).read(), 0x5EED)   # what the tests will load

    rays = fixed_ray_set(rt, po, scene, width=48, seed=7)       # 48x32 primaries + their first bounces
    extra = np.random.default_rng(3)
    d = extra.normal(size=(1024, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = extra.uniform(-12, 12, size=(1024, 3)); o[:, 1] = np.abs(o[:, 1]) * 0.2 + 0.05
    rays = np.concatenate([rays, np.concatenate([o, d], axis=1).astype(np.float32)], axis=0)
    orc = po.OracleScene(scene.desc)
    ids, hits, mg, _ = orc.closest_hit(rays, 0.001, float("inf"), margins=True)
    hit = ids >= 0
    sc = orc.scatter(rays[hit], hits[hit], seed=77)           # record i draws from the documented stream (77, i, 0): the GPU replays it
    np.savez_compressed(os.path.join(HERE, "weekend_hits.npz"), rays=rays, id=ids, t=hits["t"], n=hits["n"], p=hits["p"], front=hits["front_face"],
                        margin_second=mg["second_rel"], margin_graze=mg["graze"], margin_edge=mg["edge"], margin_tmin=mg["tmin_rel"],
                        scatter_seed=np.int64(77), scatter_o=sc["o"], scatter_d=sc["d"], scatter_atten=sc["attenuation"], scattered=sc["scattered"])
    cam = rt.default_camera(48)
    acc, st = orc.render(cam, 8192, max_depth=50, seed=12345, threads=0)
    np.save(os.path.join(HERE, "weekend_mean_48x32.npy"), (acc / 8192.0).astype(np.float64))
    print(f"rays {len(rays)} (hits {int(hit.sum())}), render {st.rays} rays in {st.seconds:.1f} s")


if __name__ == "__main__":
    main()
