"""shirley_raytracing_rs_b200 — B200-native (sm_100a) backend for the per-pixel path-tracing
loop of scottschroeder/shirley-raytracing-rs.

Layout: ``csrc/`` CUDA kernels + the C ABI (include/b200rt.h), ``host/`` the C++ mirror of
the reference's builder API (include/b200rt_host.h), ``api.py`` the thin Python face used by
tests/ and bench.py.  Importing this package loads ``libb200rt.so`` and fails loudly if it
has not been built; there is no CPU fallback.
"""
from . import _ffi  # noqa: F401  (loads the library)
from .api import *  # noqa: F401,F403
from ._ffi import B200rtError, Camera, RenderParams, Stats  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
