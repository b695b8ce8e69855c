"""Host mirror of the reference's builder API: scene factories, flatten, JSON wire format,
camera builder.  CPU only (no compute)."""
import json
import math

import numpy as np
import pytest


def test_random_scene_shape(rt, weekend):
    """src/scenes.rs:281-429: fancy ground (Rect + dielectric RectBox), 3 big spheres, ~480 small ones."""
    d = weekend.desc.contents
    F = rt._ffi
    assert 400 < d.n_prims < 490 and d.n_rects == 1 and d.n_boxes == 1 and d.n_spheres == d.n_prims - 2
    assert d.prims[0].type == F.PRIM_RECT_XZ and d.prims[1].type == F.PRIM_BOX       # create_fancy_ground, :262-278
    assert d.materials[1].kind == F.MAT_DIELECTRIC and d.materials[1].param == 1.0
    assert abs(d.rects[0].offset + 0.02) < 1e-7 and d.boxes[0].min[1] == np.float32(-0.01) and d.boxes[0].max[1] == 0.0
    big = [d.spheres[d.prims[i].index] for i in (2, 3, 4)]
    assert [s.radius for s in big] == [1.0, 1.0, 1.0] and [s.cx for s in big] == [0.0, -4.0, 4.0]
    assert d.materials[2].kind == F.MAT_DIELECTRIC and d.materials[2].param == 1.5
    assert d.materials[3].kind == F.MAT_LAMBERTIAN and d.materials[4].kind == F.MAT_METAL and d.materials[4].param == 0.0
    assert d.skybox.kind == F.SKY_ABOVE and d.n_perlin == 1
    # ground texture: checker(3, noise(1), solid .1), :256-260
    t = d.textures[d.materials[0].texture]
    assert t.kind == F.TEX_CHECKER and t.scalar == 3.0 and d.textures[t.odd].kind == F.TEX_PERLIN and d.textures[t.even].kind == F.TEX_SOLID
    kinds = [d.materials[i].kind for i in range(5, d.n_prims)]
    assert kinds.count(F.MAT_METAL) > 100 and kinds.count(F.MAT_LAMBERTIAN) > 100 and kinds.count(F.MAT_DIELECTRIC) > 20


def test_same_seed_same_scene(rt):
    a, b = rt.Scene.named("random", seed=5), rt.Scene.named("random", seed=5)
    c = rt.Scene.named("random", seed=6)
    assert a.to_json() == b.to_json() and a.to_json() != c.to_json()


def test_json_roundtrip_is_the_serde_shape(rt, weekend):
    js = json.loads(weekend.to_json())
    assert js["skybox"] == "Above"
    o0, o1, o2 = js["objects"][:3]
    assert set(o0) == {"geometry", "material"}
    assert o0["geometry"] == {"RectXZ": {"d1_min": -30.0, "d1_max": 30.0, "d2_min": -30.0, "d2_max": 30.0, "offset": -0.02}}
    assert o0["material"]["Lambertian"]["albedo"]["Checker"]["odd"] == {"Perlin": 1.0}
    assert set(o1["geometry"]["RectBox"]) == {"min", "max", "xy_sides", "yz_sides", "xz_sides"}
    assert o2["geometry"] == {"Sphere": {"center": {"vec": [0.0, 1.0, 0.0]}, "radius": 1.0}} and o2["material"] == {"Dielectric": {"ir": 1.5}}
    again = rt.Scene.from_json(weekend.to_json())
    assert again.to_json() == weekend.to_json()
    a, b = weekend.desc.contents, again.desc.contents
    assert a.n_prims == b.n_prims and a.n_textures == b.n_textures
    assert all(a.spheres[i].radius == b.spheres[i].radius for i in range(a.n_spheres))


def test_builder_api_mirrors_the_reference(rt):
    b = rt.SceneBuilder()
    b.set_skybox(rt.SkyBox.None_)
    tex = rt.TextureLoader.checker(10.0, rt.TextureLoader.solid(0.2, 0.3, 0.1), rt.TextureLoader.solid(0.9, 0.9, 0.9))
    b.add(rt.xz_rect(-30, 30, -30, 30, -0.0001), rt.Lambertian(tex))
    b.add(rt.Sphere((4, 1, 1), 1.0), rt.Lambertian(rt.TextureLoader.EarthBuiltin))
    b.add(rt.RectBox((0, 0, 0), (1, 2, 3)), rt.Metal((0.8, 0.6, 0.2), 7.0))       # fuzz clamps to 1, metal.rs:18-21
    b.add(rt.yz_rect(0, 1, 0, 1, 5), rt.DiffuseLight(rt.TextureLoader.solid(4, 4, 4)))
    b.add(rt.xy_rect(0, 1, 0, 1, 5), rt.FairyLight(rt.TextureLoader.solid(0.9, 0.9, 0.9)))   # same Solid as the checker's even child
    s = b.finalize()
    d = s.desc.contents
    F = rt._ffi
    assert d.skybox.kind == F.SKY_NONE and d.n_prims == 5 and d.n_rects == 3 and d.n_boxes == 1 and d.n_images == 1
    assert (d.images[0].width, d.images[0].height) == (1024, 512)                  # stand-in with earthmap.jpg's shape
    assert d.materials[2].param == 1.0
    # texture dedup (loader.rs:113-131): Solid(.9,.9,.9) is loaded once
    assert d.materials[4].texture == d.textures[d.materials[0].texture].even
    with pytest.raises(rt.B200rtError):
        bad = rt.SceneBuilder()
        bad.add(rt.Sphere((0, 0, 0), 1), rt.Lambertian(rt.TextureLoader.ImagePath("/no/such/file.jpg")))
        bad.finalize()
    with pytest.raises(rt.B200rtError):
        rt.Scene.from_json('{"skybox": "Sideways", "objects": []}')


def test_default_camera(rt):
    """src/scenes.rs:214-231 + camera/mod.rs:44-85."""
    cam = rt.default_camera(1200)
    assert (cam.image_width, cam.image_height) == (1200, 800)
    h = 2 * math.tan(math.radians(20.0) / 2)
    assert cam.height == pytest.approx(h, rel=1e-15) and cam.width == pytest.approx(1.5 * h, rel=1e-15)
    assert cam.lens_radius == 0.0005 and cam.focal_length == 1.0 and cam.focus_length == 10.0   # focus forced to 10, :229
    w = np.array(cam.w[:]); u = np.array(cam.u[:]); v = np.array(cam.v[:])
    np.testing.assert_allclose(w, np.array([13, 2, 3]) / math.sqrt(182), rtol=1e-15)
    assert abs(w @ u) < 1e-15 and abs(w @ v) < 1e-15 and abs(u @ v) < 1e-15
    cam169 = rt.default_camera(1920, aspect_ratio=(16, 9))
    assert cam169.image_height == 1080
    with pytest.raises(rt.B200rtError):    # Dimmensions::from_two_of_three needs exactly two
        rt.camera((0, 0, 1), (0, 0, 0), width=100, height=50, aspect_ratio=(3, 2))
    c = rt.camera((0, 0, 1), (0, 0, 0), width=100, aspect_ratio=(2, 1), aperture=None)
    assert c.lens_radius < 0 and c.focus_length == 1.0 and c.image_height == 50


def test_other_factories(rt):
    F = rt._ffi
    earth = rt.Scene.named("earth").desc.contents
    assert earth.n_prims == 2 and earth.n_images == 1 and earth.textures[earth.materials[1].texture].kind == F.TEX_IMAGE
    cornell = rt.Scene.named("cornell").desc.contents
    assert cornell.n_prims == 8 and cornell.n_boxes == 2 and cornell.skybox.kind == F.SKY_NONE and cornell.materials[2].kind == F.MAT_FAIRY_LIGHT
    demo = rt.Scene.named("demo").desc.contents
    assert demo.n_prims == 5 and demo.spheres[3].radius == np.float32(-0.4)
    scaled = rt.Scene.named("scaled", param=20).desc.contents
    assert 1500 < scaled.n_prims < 1610
    lat = rt.Scene.named("lattice", param=2).desc.contents
    assert lat.n_prims == 64


def test_checkpoint_roundtrip_and_corruption(rt, tmp_path):
    """Accumulation-buffer checkpoint (include/b200rt_host.h): exact round trip; bad files are refused."""
    rng = np.random.default_rng(4)
    acc = rng.uniform(0, 500, size=(19, 23, 4)).astype(np.float32)
    path = tmp_path / "frame.ckpt"
    rt.checkpoint_save(path, acc, samples_done=137, seed=0xDEADBEEF12345678)
    got, done, seed = rt.checkpoint_load(path)
    assert np.array_equal(got, acc) and done == 137 and seed == 0xDEADBEEF12345678
    raw = bytearray(path.read_bytes())
    raw[200] ^= 0x40                                        # flip one bit of the payload
    (tmp_path / "bad.ckpt").write_bytes(bytes(raw))
    with pytest.raises(rt.B200rtError, match="CRC"):
        rt.checkpoint_load(tmp_path / "bad.ckpt")
    (tmp_path / "short.ckpt").write_bytes(bytes(raw[:100]))
    with pytest.raises(rt.B200rtError):
        rt.checkpoint_load(tmp_path / "short.ckpt")
    (tmp_path / "junk.ckpt").write_bytes(b"PNG\x00" * 64)
    with pytest.raises(rt.B200rtError, match="not a b200rt checkpoint"):
        rt.checkpoint_load(tmp_path / "junk.ckpt")
    with pytest.raises(rt.B200rtError):
        rt.checkpoint_load(tmp_path / "missing.ckpt")
