"""Shared helpers for the parity tests: the fixed ray set of SURVEY.md §8c."""
import numpy as np


def sphere_scene(rt, spheres, material=None):
    """Tree of spheres like the reference's bbox_tree tests (bvh/bbox_tree.rs:122-227)."""
    b = rt.SceneBuilder()
    for c, r in spheres:
        b.add(rt.Sphere(c, r), material or rt.Lambertian(rt.TextureLoader.solid(0.5, 0.5, 0.5)))
    return b.finalize()


def fixed_ray_set(rt, po, scene, width, seed=7, bounce_fraction=1.0):
    """Primary rays through pixel centres of the default camera plus first-bounce rays the
    ORACLE scatters from their hits (f64, own RNG stream) — the 'fixed ray set' of §8c.
    Returns float32 [n, 6]."""
    cam = rt.default_camera(width)
    W, H = cam.image_width, cam.image_height
    xs, ys = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1)
    prim = po.camera_rays(cam, xy, seed=seed).astype(np.float32)
    o = po.OracleScene(scene.desc)
    ids, hits, _, _ = o.closest_hit(prim, 0.001, float("inf"))
    sel = np.nonzero(ids >= 0)[0]
    if bounce_fraction < 1.0:
        sel = sel[: int(len(sel) * bounce_fraction)]
    sc = o.scatter(prim[sel], hits[sel], seed=seed + 1)
    ok = sc["scattered"] != 0
    bounce = np.concatenate([sc["o"][ok], sc["d"][ok]], axis=1).astype(np.float32)
    o.close()
    return np.concatenate([prim, bounce], axis=0)


def random_rays(n, seed, origin_scale=15.0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-origin_scale, origin_scale, size=(n, 3))
    o[:, 1] = np.abs(o[:, 1]) * 0.3 + 0.05
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1).astype(np.float32)


def decidable(margins, thr=1e-4):
    """Rays whose closest hit is numerically decidable at f32 (SURVEY.md §8c check 1)."""
    return (margins["second_rel"] > thr) & (margins["graze"] > thr) & (margins["edge"] > thr * 0.1) & (margins["tmin_rel"] > thr)
