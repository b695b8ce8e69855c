"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["b200rt.h", "b200rt_host.h"])
def test_every_declared_symbol_is_exported(rt, header):
    names = declared_symbols(header)
    assert len(names) >= 9
    lib = C.CDLL(rt._ffi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding covers the whole header
    unbound = [n for n in names if n not in rt._ffi.SIGNATURES]
    assert not unbound, unbound


def test_struct_sizes(rt):
    F = rt._ffi
    assert C.sizeof(F.Sphere) == 16 and C.sizeof(F.Rect) == 32 and C.sizeof(F.Box) == 32
    assert C.sizeof(F.Material) == 32 and C.sizeof(F.Texture) == 32 and C.sizeof(F.PrimRef) == 8
    assert C.sizeof(F.Perlin) == 256 * 12 + 768
    assert C.sizeof(F.Ray) == 24 and C.sizeof(F.Hit) == 44


def test_abi_version(rt):
    assert rt._ffi.lib.b200rt_abi_version() == rt._ffi.ABI_VERSION


def test_no_cpu_fallback(rt, weekend):
    """Without a CUDA device every compute entry point returns B200RT_ECUDA with a message."""
    if rt.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rt.B200rtError) as e:
        weekend.device()
    assert e.value.code == rt._ffi.ECUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(rt.B200rtError):
        rt.fp32_peak()
    with pytest.raises(rt.B200rtError):
        rt.aabb_hit([[0, 0, 0, 1, 1, 1]], [[0, 0, 0, 1, 0, 0]])


def test_invalid_descriptions_are_rejected(rt):
    F = rt._ffi
    d = F.SceneDesc(abi_version=99)
    h = C.c_void_p()
    assert F.lib.b200rt_scene_create(C.byref(d), -1, C.byref(h)) == F.EINVAL
    assert b"abi_version" in F.lib.b200rt_last_error()
    d = F.SceneDesc(abi_version=F.ABI_VERSION, n_prims=1)   # NULL arrays
    assert F.lib.b200rt_scene_create(C.byref(d), -1, C.byref(h)) == F.EINVAL


def test_png_encoder_roundtrip(rt, tmp_path):
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    path = tmp_path / "x.png"
    rt.write_png(path, img)
    back = np.asarray(Image.open(path).convert("RGB"))
    assert np.array_equal(back, img)
    with pytest.raises(rt.B200rtError) as e:
        rt.write_png("/nonexistent-dir/x.png", img)
    assert e.value.code == rt._ffi.EIO
