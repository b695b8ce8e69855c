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


def _handmade_png(samples, depth, ctype, interlace, palette=None):
    """A PNG written chunk by chunk (filter type 0 or 1 per row): samples is (H, W, channels) of integers < 2**depth."""
    import struct
    import zlib
    H, W, ch = samples.shape

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)

    def rows(sub, flt):
        out = bytearray()
        for r in sub:
            flat = r.reshape(-1)
            if depth == 16:
                line = flat.astype(">u2").tobytes()
            elif depth == 8:
                line = flat.astype(np.uint8).tobytes()
            else:
                bits = "".join(format(int(v), f"0{depth}b") for v in flat)
                bits += "0" * (-len(bits) % 8)
                line = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
            if flt == 1:                                   # Sub filter: difference to the byte one pixel to the left
                bpp = max(1, ch * depth // 8)
                b = bytearray(line)
                for i in range(len(b) - 1, bpp - 1, -1):
                    b[i] = (b[i] - b[i - bpp]) & 0xFF
                line = bytes(b)
            out += bytes([flt]) + line
        return bytes(out)

    if interlace:
        passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
        raw = b"".join(rows(samples[y0::dy, x0::dx], k % 2) for k, (x0, y0, dx, dy) in enumerate(passes) if samples[y0::dy, x0::dx].size)
    else:
        raw = rows(samples, 1)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        png += chunk(b"PLTE", bytes(palette))
    return png + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")


@pytest.mark.parametrize("interlace", [False, True])
@pytest.mark.parametrize("size", [(1, 1), (5, 3), (8, 8), (19, 33), (64, 9)])
def test_interlaced_and_16_bit_png(rt, interlace, size):
    """Adam7-interlaced files and 16-bit samples (every colour type): `image::open` reads them, and `get_pixel` on the
    DynamicImage narrows u16 to u8 as (c + 128) / 257 (image 0.24 colour conversion)."""
    w, h = size
    rng = np.random.default_rng(w * 131 + h)
    to8 = lambda c: ((c.astype(np.uint32) + 128) // 257).astype(np.uint8)
    for ctype, ch in ((0, 1), (2, 3), (4, 2), (6, 4)):
        for depth in (8, 16):
            s = rng.integers(0, 1 << depth, size=(h, w, ch))
            got = rt.decode_image(_handmade_png(s, depth, ctype, interlace))
            c = to8(s) if depth == 16 else s.astype(np.uint8)
            want = np.repeat(c[..., :1], 3, axis=2) if ch <= 2 else c[..., :3]
            assert np.array_equal(got, want), (ctype, depth, interlace)
    for depth in (1, 2, 4):                                 # sub-byte greyscale and palette indices through the interlacer
        s = rng.integers(0, 1 << depth, size=(h, w, 1))
        got = rt.decode_image(_handmade_png(s, depth, 0, interlace))
        g = (s[..., 0] * 255 // ((1 << depth) - 1)).astype(np.uint8)
        assert np.array_equal(got, np.repeat(g[..., None], 3, axis=2)), (depth, interlace)
        pal = rng.integers(0, 256, size=(1 << depth) * 3).astype(np.uint8)
        got = rt.decode_image(_handmade_png(s, depth, 3, interlace, palette=pal))
        assert np.array_equal(got, pal.reshape(-1, 3)[s[..., 0]]), (depth, interlace)
    if not interlace and w > 1:
        # cross-check the hand-made encoder itself against PIL on an 8-bit RGB file
        s = rng.integers(0, 256, size=(h, w, 3))
        data = _handmade_png(s, 8, 2, True)
        assert np.array_equal(_want(data), s.astype(np.uint8)) and np.array_equal(rt.decode_image(data), s.astype(np.uint8))


def test_png_rejects_bad_input(rt):
    good = _png(PIL.fromarray(np.zeros((8, 8, 3), np.uint8)))
    bad = bytearray(good); bad[40] ^= 0xFF
    with pytest.raises(rt.B200rtError, match="CRC|inflate"):
        rt.decode_image(bytes(bad))
    with pytest.raises(rt.B200rtError):
        rt.decode_image(good[:30])
    img16 = PIL.fromarray((np.arange(64, dtype=np.uint16).reshape(8, 8) * 900))                 # PIL writes 16-bit greyscale
    want = ((np.arange(64, dtype=np.uint32).reshape(8, 8) * 900 + 128) // 257).astype(np.uint8)
    assert np.array_equal(rt.decode_image(_png(img16))[..., 0], want)
    with pytest.raises(rt.B200rtError, match="bit depth"):
        rt.decode_image(_handmade_png(np.zeros((2, 2, 3), int), 4, 2, False))                   # 4-bit RGB does not exist
    with pytest.raises(rt.B200rtError, match="interlace"):
        bad = bytearray(_handmade_png(np.zeros((2, 2, 1), int), 8, 0, False))
        bad[28] = 2                                                                              # interlace method 2
        import struct, zlib
        bad[29:33] = struct.pack(">I", zlib.crc32(bytes(bad[12:29])) & 0xFFFFFFFF)
        rt.decode_image(bytes(bad))


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
    q = tmp_path / "tex.gif"
    PIL.fromarray(arr).save(q)
    b2 = rt.SceneBuilder()
    b2.add(rt.Sphere((0, 0, 0), 1.0), rt.Lambertian(rt.TextureLoader.ImagePath(str(q))))
    with pytest.raises(rt.B200rtError, match="unsupported format"):
        b2.finalize()


@pytest.mark.parametrize("size", [(1, 1), (5, 3), (33, 7), (64, 16)])
def test_bmp_and_pnm_match_pil(rt, size, tmp_path):
    """`image::open` also sniffs BMP and PNM; the host mirror decodes the uncompressed forms of both."""
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    arr = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    for mode in ("RGB", "L", "P", "1", "RGBA"):
        img = PIL.fromarray(arr)
        img = img.quantize(colors=16) if mode == "P" else (img.convert("L").point(lambda v: 255 if v > 127 else 0).convert("1") if mode == "1" else img.convert(mode))
        buf = io.BytesIO(); img.save(buf, format="BMP"); data = buf.getvalue()
        want = np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))
        assert np.array_equal(rt.decode_image(data), want), ("bmp", mode)
    for mode, fmt in (("RGB", "PPM"), ("L", "PPM")):
        buf = io.BytesIO(); PIL.fromarray(arr).convert(mode).save(buf, format=fmt); data = buf.getvalue()
        assert np.array_equal(rt.decode_image(data), np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))), ("pnm", mode)
    ascii_ppm = f"P3\n# a comment\n{w} {h}\n1023\n".encode() + " ".join(str(int(v) * 4) for v in arr.reshape(-1)).encode()
    got = rt.decode_image(ascii_ppm)
    assert np.array_equal(got, ((arr.astype(np.uint32) * 4 * 255 + 511) // 1023).astype(np.uint8))
    p = tmp_path / "tex.bmp"
    PIL.fromarray(arr).save(p)
    b = rt.SceneBuilder()
    b.add(rt.Sphere((0, 0, 0), 1.0), rt.Lambertian(rt.TextureLoader.ImagePath(str(p))))
    scene = b.finalize()
    d = scene.desc.contents
    assert np.array_equal(np.ctypeslib.as_array(d.images[0].rgb8, shape=(h, w, 3)), arr)
    for bad in (b"BM" + bytes(60), b"P6\n3 3\n255\n" + bytes(5), b"P6 0 0 255 "):
        with pytest.raises(rt.B200rtError):
            rt.decode_image(bad)
