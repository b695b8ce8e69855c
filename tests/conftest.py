import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_if_needed():
    import subprocess
    need = [os.path.join(ROOT, "shirley_raytracing_rs_b200", "libb200rt.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT, "all"], stdout=subprocess.DEVNULL)


_build_if_needed()


@pytest.fixture(scope="session")
def rt():
    import shirley_raytracing_rs_b200 as m
    return m


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def weekend(rt):
    """BASELINE config 1/2 scene: src/scenes.rs random_scene (day), seeded."""
    return rt.Scene.named("random", seed=0xDEADBEEF)


@pytest.fixture(scope="session")
def gpu_required(rt):
    if rt.device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible; the product has no CPU fallback")
    return True
