"""GPU parity at BASELINE scale (pytest -m gpu): what the toy-sized tests of test_gpu_parity.py cannot see.

  * closest-hit ids / t / normals on 1e5-primitive (global-memory accessor, host SAH tree) and >= 262 144-primitive
    (device-built linear BVH, the automatic path) scenes, against the brute-force f32 mirror and the f64 oracle;
  * a converged 300x200 render at 4096 spp against two independent 4096-spp oracle runs (committed fixture);
  * RectBox against the reference's six-rect form in f64 (Cornell boxes: ids, t, normals, u, v, front_face);
  * the builders' depth guards (clustered scenes; linear BVH deeper than the stack falls back to the host builder);
  * the accumulation contract for non-finite and huge samples.
"""
import ctypes as C
import os

import numpy as np
import pytest

from common import decidable, random_rays

pytestmark = pytest.mark.gpu
INF = float("inf")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("G,n_rays,n_f64,builder", [(158, 20_000, 20_000, 0), (260, 20_000, 8_000, 1)])
def test_ids_at_baseline_scale(rt, po, gpu_required, G, n_rays, n_f64, builder):
    """BASELINE config 4 shape: `scaled` G=158 is ~1e5 spheres (tree in global memory, host SAH), G=260 is ~2.7e5
    (>= 262 144: scene_create switches to the device-built linear BVH by itself, 40-50 levels deep)."""
    s = rt.Scene.named("scaled", seed=11, param=G)
    info = s.info()
    assert info.n_prims > (90_000 if G == 158 else 262_144)
    # tree in global memory (nothing staged), one inner node fewer than leaves; the ground rect and its coat box are the up-front list
    assert info.bvh_builder == builder and info.bvh_nodes_in_smem == 0 and info.n_prims - 8 <= info.n_bvh_nodes + 1 <= info.n_prims
    rays = random_rays(n_rays, 5, origin_scale=float(G))
    ids, hits, st = rt.closest_hit(s, rays, 0.001, INF)
    want = po.closest_hit_gpu32(s.desc, rays, 0.001, INF)                 # every primitive in id order, the device's f32 arithmetic
    assert np.array_equal(ids, want["id"]), f"{(ids != want['id']).sum()} id mismatches"
    hit = ids >= 0
    assert 0.2 < hit.mean() < 0.9
    for f in ("t", "p", "n", "front_face"):
        assert np.array_equal(hits[f][hit], want[f][hit]), f
    # f64 oracle (median-split tree: the reference's own constructor is O(N^2) and degenerates at this size)
    o = po.OracleScene(s.desc, reference_topology=False)
    sub = rays[:n_f64]
    ids64, h64, mg, _ = o.closest_hit(sub, 0.001, INF, margins=True)
    keep = decidable(mg)
    assert keep.mean() > 0.99
    assert np.array_equal(ids[:n_f64][keep], ids64[keep])
    both = keep & (ids64 >= 0)
    rel = np.abs(hits["t"][:n_f64][both] - h64["t"][both]) / h64["t"][both]
    assert np.quantile(rel, 0.999) < 1e-5 and rel.max() < 5e-5
    # normals: (p - c) / r carries the f32 rounding of the hit point, eps * |origin - centre| in space, divided by a
    # radius down to 0.05 with origins hundreds of units away — the bound scales with that ratio (1e-5 at Weekend scale)
    d = s.desc.contents
    prim_index = np.array([d.prims[int(i)].index for i in ids64[both]])
    sph = np.frombuffer((rt._ffi.Sphere * d.n_spheres).from_address(C.addressof(d.spheres.contents)), dtype=np.float32).reshape(-1, 4)[prim_index]
    is_sphere = np.array([d.prims[int(i)].type == rt._ffi.PRIM_SPHERE for i in ids64[both]])
    ratio = np.linalg.norm(sub[both][:, :3] - sph[:, :3], axis=1) / np.abs(sph[:, 3])
    nerr = np.abs(hits["n"][:n_f64][both] - h64["n"][both]).max(axis=1)
    assert np.all(nerr[is_sphere] <= 4 * 1.2e-7 * ratio[is_sphere] + 1e-6), (nerr[is_sphere] / (1.2e-7 * ratio[is_sphere] + 1e-9)).max()
    assert np.all(nerr[~is_sphere] == 0)                                                   # axis-aligned faces
    # the render kernel on the same scene: deterministic, tile shards reassemble the frame bit for bit
    cam = rt.camera((0.9 * G, 0.18 * G + 2, 0.35 * G), (0, 0, 0), vfov=30, aperture=0.001, width=256, aspect_ratio=(16, 9), focus_length=10.0)
    full, stf = rt.render(s, cam, samples=4, seed=3)
    parts = [rt.render(s, cam, samples=4, seed=3, shard=(4, k))[0] for k in range(4)]
    assert np.array_equal(sum(parts), full) and stf.rays > stf.paths


def test_scaled_scene_node_count(rt, gpu_required):
    """One inner node fewer than leaves; scene-spanning primitives live in the up-front list, not in the tree."""
    s = rt.Scene.named("scaled", seed=11, param=40)
    info = s.info()
    assert info.n_prims - 8 <= info.n_bvh_nodes + 1 <= info.n_prims


def test_converged_render_vs_golden(rt, po, gpu_required):
    """SURVEY.md §8c check 3 at >= 4096 vs >= 4096 spp: the linear (pre-gamma) mean of a 300x200 frame against the
    committed means of two independent f64 oracle runs (tests/golden/make_golden_converged.py).  An unbiased 4096-spp
    estimate differs from either oracle run by the oracle-vs-oracle floor (= sqrt 2 x the standard error of one run);
    the bound is 1.25 x that floor, well inside the survey's "2 x the Monte-Carlo error"."""
    g = np.load(os.path.join(HERE, "golden", "weekend_mean_300x200_4096spp.npz"))
    spp = int(g["spp"])
    assert spp >= 4096
    scene = rt.Scene.from_json(open(os.path.join(HERE, "golden", "weekend_scene.json")).read(), 0x5EED)
    cam = rt.default_camera(300)
    acc, st = rt.render(scene, cam, samples=spp, seed=777)
    m = acc[..., :3].astype(np.float64) / spp
    ma, mb, floor = g["mean_a"].astype(np.float64), g["mean_b"].astype(np.float64), float(g["floor"])
    assert m.shape == ma.shape == (200, 300, 3)
    rmse = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    ra, rb = rmse(m, ma), rmse(m, mb)
    assert abs(rmse(ma, mb) - floor) < 1e-6
    assert ra < 1.25 * floor and rb < 1.25 * floor, (ra, rb, floor)
    # against the 8192-spp average of the two oracle runs the GPU frame must be CLOSER than either run is to the other
    assert rmse(m, 0.5 * (ma + mb)) < floor, (rmse(m, 0.5 * (ma + mb)), floor)
    # no bias hiding under the noise: image-wide and per-channel means, and coarse 10x10 blocks (noise averages out 10x there)
    assert abs(m.mean() - 0.5 * (ma.mean() + mb.mean())) < 6e-4
    blk = lambda x: x.reshape(20, 10, 30, 10, 3).mean(axis=(1, 3))
    assert rmse(blk(m), blk(0.5 * (ma + mb))) < 0.25 * floor
    assert abs(st.rays / st.paths - float(g["segments"])) < 2e-3 * float(g["segments"])


def test_cornell_boxes_vs_six_rect_reference(rt, po, gpu_required):
    """RectBox::hit (geometry/rect.rs:132-156) is six rect tests in the reference (xy, yz, xz pairs, a later face
    replacing an earlier one at equal t); the device tests three slabs.  Against the f64 oracle (which restates the
    six-rect form): ids, t, normals, u, v and front_face on every decidable ray."""
    s = rt.Scene.named("cornell", seed=11)
    d = s.desc.contents
    assert d.n_boxes >= 2
    rays = random_rays(300_000, 6, origin_scale=500.0)
    rays[:, :3] = np.abs(rays[:, :3]) % 555.0
    # aim half of the rays at the two boxes so that most of them hit one
    rng = np.random.default_rng(9)
    k = len(rays) // 2
    tgt = np.stack([rng.uniform(100, 450, k), rng.uniform(0, 330, k), rng.uniform(60, 460, k)], axis=1)
    dirs = tgt - rays[:k, :3]
    rays[:k, 3:] = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    ids, hits, _ = rt.closest_hit(s, rays, 0.001, INF)
    o = po.OracleScene(s.desc)
    ids64, h64, mg, _ = o.closest_hit(rays, 0.001, INF, margins=True)
    keep = decidable(mg)
    assert keep.mean() > 0.97                         # rays aimed at the boxes graze edges more often than random ones
    assert np.array_equal(ids[keep], ids64[keep]), f"{(ids[keep] != ids64[keep]).sum()} mismatches"
    F = rt._ffi
    is_box = np.array([d.prims[int(i)].type == F.PRIM_BOX if i >= 0 else False for i in ids64])
    on_box = keep & is_box
    assert on_box.sum() > 20_000
    for sel, what in ((on_box, "box"), (keep & (ids64 >= 0) & ~is_box, "rect")):
        rel = np.abs(hits["t"][sel] - h64["t"][sel]) / h64["t"][sel]
        assert np.quantile(rel, 0.999) < 1e-5 and rel.max() < 5e-5, what
        assert np.array_equal(hits["front_face"][sel], h64["front_face"][sel]), what
        # axis-aligned faces: the normal is exact unless the ray meets an edge within rounding (two faces at ~equal t)
        same_n = np.all(hits["n"][sel] == h64["n"][sel].astype(np.float32), axis=1)
        assert same_n.mean() > 0.9995, (what, same_n.mean())
        ok = sel.copy(); ok[sel] = same_n
        assert np.quantile(np.abs(hits["u"][ok] - h64["u"][ok]), 0.999) < 1e-4 and np.quantile(np.abs(hits["v"][ok] - h64["v"][ok]), 0.999) < 1e-4, what
        assert np.abs(hits["p"][ok] - h64["p"][ok]).max() < 2e-3, what       # coordinates ~555: 1e-5 relative


def _clustered(rt, n):
    """Spheres whose sizes and positions grow geometrically: a surface-area split peels one primitive per level."""
    b = rt.SceneBuilder()
    mat = rt.Lambertian(rt.TextureLoader.solid(0.5, 0.5, 0.5))
    x = 1.0
    for k in range(n):
        b.add(rt.Sphere((x, 0.0, 0.0), 0.2 * x), mat)
        x *= 1.03
    return b.finalize()


def test_host_builder_depth_guard(rt, po, gpu_required):
    """A strongly clustered scene: the SAH builder must stay inside the 60-level traversal stack (median splits near
    the limit) instead of failing with B200RT_ESTACK, and still find the right hits."""
    s = _clustered(rt, 400)
    info = s.info()
    assert info.bvh_builder == 0 and info.bvh_depth <= 60
    rng = np.random.default_rng(4)
    o = np.stack([rng.uniform(0, 1.03 ** 400, 20_000) * rng.uniform(0, 1, 20_000) ** 8, rng.uniform(-1, 1, 20_000), rng.uniform(-1, 1, 20_000)], axis=1)
    dd = rng.normal(size=(20_000, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    rays = np.concatenate([o, dd], axis=1).astype(np.float32)
    ids, hits, _ = rt.closest_hit(s, rays, 0.001, INF)
    want = po.closest_hit_gpu32(s.desc, rays, 0.001, INF)
    assert np.array_equal(ids, want["id"]) and (ids >= 0).mean() > 0.05


def test_device_builder_falls_back_when_too_deep(rt, po, gpu_required, monkeypatch):
    """A linear BVH deeper than the traversal stack is rebuilt by the host builder instead of failing the scene
    (forced here by lowering the limit: B200RT_LBVH_MAX_DEPTH is a test hook)."""
    rays = random_rays(50_000, 9, origin_scale=45.0)
    monkeypatch.setenv("B200RT_BUILDER", "lbvh")
    s1 = rt.Scene.named("scaled", seed=11, param=40)
    assert s1.info().bvh_builder == 1
    ids1, _, _ = rt.closest_hit(s1, rays, 0.001, INF)
    monkeypatch.setenv("B200RT_LBVH_MAX_DEPTH", "5")
    s2 = rt.Scene.named("scaled", seed=11, param=40)
    info = s2.info()
    assert info.bvh_builder == 0 and 5 < info.bvh_depth <= 60
    ids2, _, _ = rt.closest_hit(s2, rays, 0.001, INF)
    assert np.array_equal(ids1, ids2)
    # the same through a genuinely clustered scene on the device builder: whatever depth it reaches, the scene is usable
    monkeypatch.delenv("B200RT_LBVH_MAX_DEPTH")
    s3 = _clustered(rt, 400)
    assert s3.info().bvh_depth <= 60
    r3 = random_rays(5_000, 2, origin_scale=50.0)
    assert np.array_equal(rt.closest_hit(s3, r3, 0.001, INF)[0], po.closest_hit_gpu32(s3.desc, r3, 0.001, INF)["id"])


def test_non_finite_and_huge_samples(rt, gpu_required):
    """Accumulation contract: a NaN / infinite sample contributes nothing (the reference would carry
    the NaN into the pixel and write 0); a huge finite sample is clamped to 2e9 / samples per channel, so the pixel's
    fixed-point sum cannot wrap: the pixel saturates instead of going negative or dark."""
    cam = rt.camera((0, 0, 5), (0, 0, 0), vfov=40, aperture=None, width=32, aspect_ratio=(1, 1), focus_length=10.0)
    for value, expect in ((float("nan"), "dropped"), (float("inf"), "dropped"), (3.0e38, "clamped"), (1.0e12, "clamped"), (4.0, "plain")):
        b = rt.SceneBuilder().set_skybox(rt.SkyBox.Flat((0.25, 0.25, 0.25)))
        b.add(rt.Sphere((0, 0, 0), 1.0), rt.DiffuseLight(rt.TextureLoader.solid(value, value, value)))
        s = b.finalize()
        spp = 16
        acc, st = rt.render(s, cam, samples=spp, seed=1)
        assert np.all(np.isfinite(acc)) and np.all(acc[..., 3] == spp)
        centre, corner = acc[16, 16, :3], acc[0, 0, :3]
        np.testing.assert_allclose(corner, 0.25 * spp, rtol=1e-6)           # background pixels are untouched
        if expect == "dropped":
            assert np.all(centre == 0.0)                                    # every sample of this pixel hit the emitter
        elif expect == "clamped":
            np.testing.assert_allclose(centre, 2.0e9, rtol=1e-5)            # 16 x (2e9 / 16)
            assert np.all(rt.resolve_rgb8(acc, samples=spp)[15, 16] == 255)
        else:
            np.testing.assert_allclose(centre, value * spp, rtol=1e-6)


def test_flag_wait_time_out_poisons_the_frame(rt, gpu_required):
    """b200rt_peer_wait_device gives up on a peer that never signals; the fused resolve enqueued behind it must then
    store NOTHING, and b200rt_peer_timed_out reports the rank once (reading clears it)."""
    import ctypes as C
    import torch
    F = rt._ffi
    H, W = 16, 24
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    acc = torch.ones((H, W, 4), dtype=torch.float32, device="cuda")
    out = torch.full((H, W, 3), 7, dtype=torch.uint8, device="cuda")
    ptrs = (C.c_void_p * 2)(acc.data_ptr(), acc.data_ptr())
    fl = (C.c_void_p * 2)(flags.data_ptr(), flags.data_ptr())
    # rank 0 of 2 signals itself; "rank 1" never does
    F.check(F.lib.b200rt_peer_signal_device(fl, 1, 0, 0, 1, None))
    F.check(F.lib.b200rt_peer_wait_device(flags.data_ptr(), 2, 0, 1, 50, None))           # 50 ms
    F.check(F.lib.b200rt_resolve_peers_rgb8_device(ptrs, 2, W, H, 2, 0, 0, out.data_ptr(), flags.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.all(out == 7)                                                            # nothing was stored
    t = C.c_uint32()
    F.check(F.lib.b200rt_peer_timed_out(flags.data_ptr(), C.byref(t)))
    assert t.value == 2                                                                   # 1 + the missing rank
    F.check(F.lib.b200rt_peer_timed_out(flags.data_ptr(), C.byref(t)))
    assert t.value == 0                                                                   # reported once
    # with the flag clear the same call resolves the frame
    F.check(F.lib.b200rt_resolve_peers_rgb8_device(ptrs, 2, W, H, 2, 0, 0, out.data_ptr(), flags.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.all(out == 255)


def test_multi_handle_on_one_device(rt, weekend, gpu_required):
    """b200rt_multi_create / _render_rgb8 / _destroy with a single device: the frame-loop entry gives the same bytes as
    render_scene, frame after frame (buffers and streams are reused), and a bad device list is refused."""
    import ctypes as C
    F = rt._ffi
    cam = rt.default_camera(200)
    H, W = cam.image_height, cam.image_width
    devs = (C.c_int * 1)(0)
    mh = C.c_void_p()
    F.check(F.lib.b200rt_multi_create(devs, 1, C.byref(mh)))
    try:
        for seed in (3, 4, 3):
            want, st1 = rt.render_scene(weekend, cam, samples=6, max_reflect=50, output=None, seed=seed)
            rgb = np.empty((H, W, 3), dtype=np.uint8)
            p = F.RenderParams(samples=6, max_depth=50, seed=seed, device=-1)
            st = F.Stats()
            F.check(F.lib.b200rt_multi_render_rgb8(mh, weekend.desc, C.byref(cam), C.byref(p), rgb.ctypes.data, C.byref(st)))
            assert np.array_equal(rgb, want) and st.rays == st1.rays
        cam2 = rt.default_camera(320)                                                       # a larger frame: buffers grow
        rgb = np.empty((cam2.image_height, cam2.image_width, 3), dtype=np.uint8)
        p = F.RenderParams(samples=2, max_depth=50, seed=1, device=-1)
        F.check(F.lib.b200rt_multi_render_rgb8(mh, weekend.desc, C.byref(cam2), C.byref(p), rgb.ctypes.data, None))
        assert np.array_equal(rgb, rt.render_scene(weekend, cam2, samples=2, output=None, seed=1)[0])
    finally:
        F.lib.b200rt_multi_destroy(mh)
    bad = (C.c_int * 2)(0, 0)
    assert F.lib.b200rt_multi_create(bad, 2, C.byref(mh)) == F.EINVAL
    far = (C.c_int * 1)(99)
    assert F.lib.b200rt_multi_create(far, 1, C.byref(mh)) == F.EINVAL


def test_earth_scene_uses_the_reference_asset(rt, po, gpu_required):
    """BASELINE config 3: EarthBuiltin decodes the reference's assets/earthmap.jpg (shipped next to the library) —
    the decoded bytes are the ones PIL / OpenCV produce (sha256 in SURVEY.md §8c) — and those texels are what the device
    samples (image_texture.rs:34-56, nearest texel below, v flipped)."""
    import hashlib
    earth = rt.Scene.named("earth")
    d = earth.desc.contents
    assert d.n_images == 1 and (d.images[0].width, d.images[0].height) == (1024, 512)
    texels = np.ctypeslib.as_array(d.images[0].rgb8, shape=(512, 1024, 3))
    assert hashlib.sha256(texels.tobytes()).hexdigest() == "a8cdc92a168d554ddc693785d31f5e251063724f44099571d7fbce3b43d44c45"
    tex = d.materials[1].texture
    rng = np.random.default_rng(8)
    n = 200_000
    uvp = np.zeros((n, 5), dtype=np.float32)
    uvp[:, :2] = rng.uniform(-0.1, 1.1, size=(n, 2))
    got = rt.texture_value(earth, tex, uvp)
    want = po.OracleScene(earth.desc).texture_value(tex, uvp.astype(np.float64))
    assert (np.abs(got - want).max(axis=1) > 1e-6).mean() < 1e-3            # u * (W - 1) straddling an integer in f32
    # texel centres: exact bytes / 255
    i, j = rng.integers(0, 1024, 5000), rng.integers(0, 512, 5000)
    uvp = np.zeros((5000, 5), dtype=np.float32)
    uvp[:, 0] = (i + 0.5) / 1023.0; uvp[:, 1] = 1.0 - (j + 0.5) / 511.0
    ok = (i < 1023) & (j < 511)
    got = rt.texture_value(earth, tex, uvp)
    np.testing.assert_allclose(got[ok], texels[j[ok], i[ok]].astype(np.float32) / 255.0, rtol=1e-6)
    # and a render of the scene is within Monte-Carlo noise of the oracle's on the same texture
    cam = rt.default_camera(120, aspect_ratio=(16, 9))
    spp = 256
    o = po.OracleScene(earth.desc)
    a, _ = o.render(cam, spp, seed=1)
    b, _ = o.render(cam, spp, seed=2)
    g, _ = rt.render(earth, cam, samples=spp, seed=3)
    rmse = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rmse(a / spp, b / spp)
    assert 0.5 * (rmse(g[..., :3] / spp, a / spp) + rmse(g[..., :3] / spp, b / spp)) < 1.25 * floor + 1e-4


def test_concurrent_renders_from_host_threads(rt, weekend, gpu_required):
    """SURVEY §8b threading: a scene handle is immutable after create and several host threads may render from it at
    once (rayon workers share `&Frame`, src/main.rs:118-125), each call on its own stream and scratch; scenes may be
    created and destroyed meanwhile.  Every concurrent result equals the sequential one bit for bit."""
    import threading
    cam = rt.default_camera(150)
    seeds = list(range(40, 48))
    want = {s: rt.render(weekend, cam, samples=6, seed=s)[0] for s in seeds}
    got, errors = {}, []

    def worker(s):
        try:
            for _ in range(3):
                got[s] = rt.render(weekend, cam, samples=6, seed=s)[0]
        except Exception as e:                                   # noqa: BLE001 - reported below
            errors.append((s, repr(e)))

    def churn():
        try:
            for k in range(12):
                # scenes of different sizes: their launches ask for different amounts of shared memory (the per-function
                # MaxDynamicSharedMemorySize attribute once raced between such threads)
                other = rt.Scene.named(("cornell", "random", "perlin")[k % 3], seed=100 + k)
                a, _ = rt.render(other, cam, samples=2, seed=1)
                assert np.isfinite(a).all()
                del other
        except Exception as e:                                   # noqa: BLE001
            errors.append(("churn", repr(e)))

    threads = [threading.Thread(target=worker, args=(s,)) for s in seeds] + [threading.Thread(target=churn)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    assert all(not t.is_alive() for t in threads)
    for s in seeds:
        assert np.array_equal(got[s], want[s]), s
