"""Host-side PNG decoder (shirley_raytracing_rs_b200/host/png_decoder.cpp) for TextureLoader::ImagePath
(image_texture.rs:23-26 `image::open`): identical RGB bytes to PIL for every supported layout."""
import io

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")


def _png(img, **kw):
    buf = io.BytesIO()
    img.save(buf, format="PNG", **kw)
    return buf.getvalue()


def _want(data):
    return np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P", "1"])
@pytest.mark.parametrize("size", [(37, 23), (1, 1), (64, 5), (3, 70)])
def test_png_matches_pil(rt, mode, size):
    rng = np.random.default_rng(hash((mode, size)) % 2**32)
    w, h = size
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 7 + y * 3) % 256, (x * x + y) % 256, rng.integers(0, 256, size=(h, w))], axis=-1).astype(np.uint8)
    img = PIL.fromarray(base, "RGB")
    if mode == "RGBA":
        img.putalpha(PIL.fromarray(rng.integers(0, 256, size=(h, w), dtype=np.uint8)))
    elif mode == "P":
        img = img.quantize(colors=min(64, max(2, w * h)))
    elif mode == "1":
        img = img.convert("L").point(lambda v: 255 if v > 127 else 0).convert("1")
    elif mode != "RGB":
        img = img.convert(mode)
    for kw in ({}, {"compress_level": 1}, {"optimize": True}):
        data = _png(img, **kw)
        got = rt.decode_image(data)
        want = _want(data) if mode not in ("RGBA", "LA") else np.asarray(PIL.open(io.BytesIO(data)).convert("RGBA" if mode == "RGBA" else "LA").convert("RGBA"))[..., :3]
        assert got.shape == want.shape and np.array_equal(got, want), (mode, size, kw)


def test_low_bit_depth_grey_and_palette(rt):
    """1/2/4-bit samples (PNG spec 7.2): written by hand through PIL's `bits` option."""
    g = (np.arange(16 * 9).reshape(9, 16) % 4).astype(np.uint8)
    img = PIL.fromarray(g, "P")
    img.putpalette([0, 0, 0, 255, 0, 0, 0, 255, 0, 0, 0, 255] + [0] * (252 * 3))
    for bits in (2, 4, 8):
        data = _png(img, bits=bits)
        assert np.array_equal(rt.decode_image(data), _want(data)), bits


def test_png_rejects_bad_input(rt):
    good = _png(PIL.fromarray(np.zeros((8, 8, 3), np.uint8)))
    bad = bytearray(good); bad[40] ^= 0xFF
    with pytest.raises(rt.B200rtError, match="CRC|inflate"):
        rt.decode_image(bytes(bad))
    with pytest.raises(rt.B200rtError):
        rt.decode_image(good[:30])
    inter = io.BytesIO()
    img16 = PIL.fromarray((np.arange(64, dtype=np.uint16).reshape(8, 8) * 900))
    with pytest.raises(rt.B200rtError, match="16-bit"):
        rt.decode_image(_png(img16))


def test_png_image_path_texture(rt, tmp_path):
    arr = (np.random.default_rng(1).integers(0, 256, size=(12, 20, 3))).astype(np.uint8)
    p = tmp_path / "tex.png"
    PIL.fromarray(arr).save(p)
    b = rt.SceneBuilder()
    b.add(rt.Sphere((0, 0, 0), 1.0), rt.Lambertian(rt.TextureLoader.ImagePath(str(p))))
    scene = b.finalize()          # keep the Scene alive: desc points into it
    d = scene.desc.contents
    assert d.n_images == 1 and (d.images[0].width, d.images[0].height) == (20, 12)
    assert np.array_equal(np.ctypeslib.as_array(d.images[0].rgb8, shape=(12, 20, 3)), arr)
    q = tmp_path / "tex.bmp"
    PIL.fromarray(arr).save(q)
    b2 = rt.SceneBuilder()
    b2.add(rt.Sphere((0, 0, 0), 1.0), rt.Lambertian(rt.TextureLoader.ImagePath(str(q))))
    with pytest.raises(rt.B200rtError, match="unsupported format"):
        b2.finalize()
