"""The host binned-SAH BVH builder (csrc/bvh_build.hpp) on adversarial inputs, compiled into a small native harness:
depth stays inside the 60-level device traversal stack (median splits take over near the limit), each primitive is a
leaf exactly once, child boxes nest.  CPU only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_builder_depth_guard_and_tree_invariants(tmp_path):
    exe = str(tmp_path / "bvh_build_check")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-o", exe, os.path.join(HERE, "native", "bvh_build_check.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" ok") == 5, r.stdout
