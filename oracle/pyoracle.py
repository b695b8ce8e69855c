"""ctypes face of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY — see oracle.hpp).

Takes the same B200rtSceneDesc the product consumes, so both sides see identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


class OracleStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("depth_exhausted", C.c_uint64), ("pops", C.c_uint64),
                ("box_hits", C.c_uint64), ("leaf_tests", C.c_uint64), ("max_stack", C.c_uint64), ("seconds", C.c_double),
                ("threads", C.c_int32), ("_pad", C.c_int32)]


HIT_DTYPE = np.dtype([("t", "<f8"), ("p", "<f8", 3), ("n", "<f8", 3), ("u", "<f8"), ("v", "<f8"), ("front_face", "<i4"), ("id", "<i4")])
MARGIN_DTYPE = np.dtype([("second_rel", "<f8"), ("graze", "<f8"), ("edge", "<f8"), ("tmin_rel", "<f8")])
SCATTER_DTYPE = np.dtype([("o", "<f8", 3), ("d", "<f8", 3), ("attenuation", "<f8", 3), ("emitted", "<f8", 3), ("scattered", "<i4"), ("draws", "<u4")])
GPU_HIT_DTYPE = np.dtype([("t", "<f4"), ("p", "<f4", 3), ("n", "<f4", 3), ("u", "<f4"), ("v", "<f4"), ("front_face", "<i4"), ("id", "<i4")])


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run `make oracle`")
    lib = C.CDLL(LIB_PATH)
    vp, sz, dbl, u64, u32, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_uint64, C.c_uint32, C.c_int
    sig = {
        "oracle_scene_create": (vp, [vp, i32, i32]),
        "oracle_scene_destroy": (None, [vp]),
        "oracle_tree_info": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "oracle_closest_hit": (i32, [vp, vp, sz, dbl, dbl, vp, vp, vp, C.POINTER(OracleStats)]),
        "oracle_closest_hit_f64": (i32, [vp, vp, sz, dbl, dbl, vp, vp]),
        "oracle_closest_hit_gpu32": (i32, [vp, vp, sz, C.c_float, C.c_float, vp]),
        "oracle_scatter": (i32, [vp, vp, vp, sz, u64, vp, sz, vp]),
        "oracle_camera_rays": (i32, [vp, i32, vp, sz, u64, vp]),
        "oracle_texture_value": (i32, [vp, C.c_int32, vp, sz, vp]),
        "oracle_render": (i32, [vp, vp, vp, vp, C.POINTER(OracleStats), i32]),
        "oracle_resolve_rgb8": (i32, [vp, u32, u32, u32, vp]),
        "oracle_rng_uniforms": (i32, [u64, u32, u32, sz, vp]),
        "oracle_aabb_hit2": (i32, [vp, vp, dbl, dbl]),
        "oracle_aabb_hit": (i32, [vp, vp, dbl, dbl]),
        "oracle_surrounding_box": (None, [vp, vp, vp]),
        "oracle_fmin": (dbl, [dbl, dbl]),
        "oracle_fmax": (dbl, [dbl, dbl]),
        "oracle_sizeof_tree_node_f64": (u64, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def _rays(rays):
    r = np.ascontiguousarray(rays, dtype=np.float32)
    assert r.ndim == 2 and r.shape[1] == 6
    return r


class OracleScene:
    """desc_ptr: a ctypes pointer to B200rtSceneDesc (e.g. shirley_raytracing_rs_b200.Scene.desc)."""

    def __init__(self, desc_ptr, reference_topology: bool = True, precision: int = 64):
        self._desc = desc_ptr
        self.precision = precision
        self._h = C.c_void_p(lib.oracle_scene_create(C.cast(desc_ptr, C.c_void_p), int(reference_topology), precision))
        if not self._h:
            raise RuntimeError("oracle_scene_create failed")

    def close(self):
        if self._h:
            lib.oracle_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree_info(self):
        n, d = C.c_uint64(), C.c_uint64()
        lib.oracle_tree_info(self._h, C.byref(n), C.byref(d))
        return n.value, d.value

    def closest_hit(self, rays, t_min=0.001, t_max=float("inf"), margins=False):
        r = _rays(rays)
        n = r.shape[0]
        ids = np.empty(n, dtype=np.int32)
        hits = np.zeros(n, dtype=HIT_DTYPE)
        mg = np.zeros(n, dtype=MARGIN_DTYPE) if margins else None
        st = OracleStats()
        rc = lib.oracle_closest_hit(self._h, r.ctypes.data, n, t_min, t_max, ids.ctypes.data, hits.ctypes.data,
                                    mg.ctypes.data if margins else None, C.byref(st))
        assert rc == 0
        return ids, hits, mg, st

    def closest_hit_f64(self, rays, t_min=0.001, t_max=float("inf")):
        """Rays as f64 (the reference's precision): exact replay of the reference's unit-test values."""
        r = np.ascontiguousarray(rays, dtype=np.float64)
        assert r.ndim == 2 and r.shape[1] == 6
        ids = np.empty(r.shape[0], dtype=np.int32)
        hits = np.zeros(r.shape[0], dtype=HIT_DTYPE)
        rc = lib.oracle_closest_hit_f64(self._h, r.ctypes.data, r.shape[0], t_min, t_max, ids.ctypes.data, hits.ctypes.data)
        assert rc == 0
        return ids, hits

    def scatter(self, rays, hits, seed=0, injected=None):
        r = _rays(rays)
        h = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = np.zeros(r.shape[0], dtype=SCATTER_DTYPE)
        inj, stride = None, 0
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float64)
            inj, stride = injected.ctypes.data, injected.shape[1]
        rc = lib.oracle_scatter(self._h, r.ctypes.data, h.ctypes.data, r.shape[0], seed, inj, stride, out.ctypes.data)
        assert rc == 0
        return out

    def texture_value(self, tex, uvp):
        q = np.ascontiguousarray(uvp, dtype=np.float64)
        out = np.empty((q.shape[0], 3), dtype=np.float64)
        rc = lib.oracle_texture_value(self._h, tex, q.ctypes.data, q.shape[0], out.ctypes.data)
        assert rc == 0
        return out

    def render(self, cam, samples, max_depth=50, seed=0, sample_offset=0, rows=(0, 0), threads=0):
        """cam: shirley_raytracing_rs_b200.Camera (ctypes). Returns (accum[H,W,3] f64 sums, OracleStats)."""
        from shirley_raytracing_rs_b200 import _ffi as F
        H, W = cam.image_height, cam.image_width
        accum = np.zeros((H, W, 3), dtype=np.float64)
        p = F.RenderParams(samples=samples, sample_offset=sample_offset, max_depth=max_depth, seed=seed,
                           row_begin=rows[0], row_end=rows[1], device=-1)
        st = OracleStats()
        rc = lib.oracle_render(self._h, C.byref(cam), C.byref(p), accum.ctypes.data, C.byref(st), threads)
        assert rc == 0
        return accum, st


def closest_hit_gpu32(desc_ptr, rays, t_min=0.001, t_max=float("inf")):
    r = _rays(rays)
    hits = np.zeros(r.shape[0], dtype=GPU_HIT_DTYPE)
    rc = lib.oracle_closest_hit_gpu32(C.cast(desc_ptr, C.c_void_p), r.ctypes.data, r.shape[0], t_min, t_max, hits.ctypes.data)
    assert rc == 0
    return hits


def camera_rays(cam, xy, seed=0, precision=64):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    out = np.empty((xy.shape[0], 6), dtype=np.float64)
    rc = lib.oracle_camera_rays(C.byref(cam), precision, xy.ctypes.data, xy.shape[0], seed, out.ctypes.data)
    assert rc == 0
    return out


def resolve_rgb8(accum3, samples):
    a = np.ascontiguousarray(accum3, dtype=np.float64)
    H, W, _ = a.shape
    out = np.empty((H, W, 3), dtype=np.uint8)
    rc = lib.oracle_resolve_rgb8(a.ctypes.data, W, H, samples, out.ctypes.data)
    assert rc == 0
    return out


def rng_uniforms(seed, a, b, n):
    out = np.empty(n, dtype=np.float64)
    lib.oracle_rng_uniforms(seed, a, b, n, out.ctypes.data)
    return out


def _d6(v):
    return (C.c_double * 6)(*[float(x) for x in v])


def aabb_hit2(box6, ray6, t_min, t_max):
    return bool(lib.oracle_aabb_hit2(_d6(box6), _d6(ray6), t_min, t_max))


def aabb_hit(box6, ray6, t_min, t_max):
    return bool(lib.oracle_aabb_hit(_d6(box6), _d6(ray6), t_min, t_max))


def surrounding_box(a6, b6):
    out = (C.c_double * 6)()
    lib.oracle_surrounding_box(_d6(a6), _d6(b6), out)
    return list(out)


fmin = lib.oracle_fmin
fmax = lib.oracle_fmax
sizeof_tree_node_f64 = lib.oracle_sizeof_tree_node_f64
