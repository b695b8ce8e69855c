"""`ray-cli` front end (src/main.rs + src/argparse.rs + src/scenes.rs of the reference)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ray-cli")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, cwd=ROOT)


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", ROOT, "cli"], stdout=subprocess.DEVNULL)


def test_test_subcommand_is_a_stub():
    r = run("test")                                   # src/main.rs:60-63
    assert r.returncode == 0 and "nothing to test" in r.stderr


def test_bad_arguments():
    assert run().returncode == 2
    assert run("render", "bogus").returncode == 2
    assert run("render", "saved").returncode == 2    # SCENE_INPUT is required
    assert run("render", "random", "--camera-aspect-ratio", "4x3").returncode == 2


def test_scene_output_is_written_before_rendering(tmp_path, rt):
    """render_random writes --scene-output before finalize/render (src/scenes.rs:140-146); the
    file is the serde shape and `render saved` reads it back."""
    out = tmp_path / "scene.json"
    r = run("render", "random", "--seed", "7", "--scene-output", str(out), "-w", "60", "-s", "1", "-o", str(tmp_path / "o.png"))
    js = json.loads(out.read_text())
    assert js["skybox"] == "Above" and len(js["objects"]) > 400
    night = tmp_path / "night.json"
    run("render", "random", "--night", "--seed", "7", "--scene-output", str(night), "-w", "60", "-s", "1", "-o", str(tmp_path / "o.png"))
    assert json.loads(night.read_text())["skybox"] == "None"
    if rt.device_count() == 0:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and "unrecoverable ray-cli failure" in r.stderr
    assert rt.Scene.from_json(out.read_text()).desc.contents.n_prims == len(js["objects"])


@pytest.mark.gpu
def test_render_end_to_end(tmp_path, gpu_required):
    from PIL import Image
    png = tmp_path / "out.png"
    scene = tmp_path / "scene.json"
    r = run("-v", "render", "random", "--seed", "3", "-w", "150", "-s", "4", "-o", str(png), "--scene-output", str(scene))
    assert r.returncode == 0, r.stderr
    img = np.asarray(Image.open(png).convert("RGB"))
    assert img.shape == (100, 150, 3) and "Mrays/s" in r.stderr
    png2 = tmp_path / "saved.png"
    r = run("render", "saved", str(scene), "--seed", "3", "-w", "150", "-s", "4", "-o", str(png2))
    assert r.returncode == 0, r.stderr
    # same scene + same seed, but Perlin tables are redrawn from the seed the same way: identical image
    assert np.array_equal(np.asarray(Image.open(png2).convert("RGB")), img)
    r = run("render", "cornell", "-w", "64", "-s", "0", "-o", str(tmp_path / "c.png"))     # samples 0 -> 1 with a warning
    assert r.returncode == 0 and "samples set to 0" in r.stderr
    assert Image.open(tmp_path / "c.png").size == (64, 64)
    r = run("render", "earth", "--camera-aspect-ratio", "std16x9", "-w", "160", "-s", "2", "-o", "/nonexistent-dir/x.png")
    assert r.returncode == 1 and "cannot write" in r.stderr


@pytest.mark.gpu
def test_checkpoint_resume_equals_one_run(tmp_path, gpu_required):
    """--checkpoint: 3 + 5 samples in two invocations == 8 samples in one (same sample sequence,
    fixed-point tile sums: identical RGB bytes)."""
    from PIL import Image
    ck = tmp_path / "acc.ckpt"
    a, b, full = tmp_path / "a.png", tmp_path / "b.png", tmp_path / "full.png"
    common = ("render", "random", "--seed", "11", "-w", "120")
    r = run(*common, "-s", "3", "-o", str(a), "--checkpoint", str(ck)); assert r.returncode == 0, r.stderr
    r = run("-v", *common, "-s", "5", "-o", str(b), "--checkpoint", str(ck)); assert r.returncode == 0, r.stderr
    assert "resuming" in r.stderr and "3 samples done" in r.stderr
    r = run(*common, "-s", "8", "-o", str(full)); assert r.returncode == 0, r.stderr
    got, want = np.asarray(Image.open(b).convert("RGB")).astype(int), np.asarray(Image.open(full).convert("RGB")).astype(int)
    assert np.abs(got - want).max() <= 1 and (got != want).mean() < 1e-3          # f32 sum of two launches vs one: last-bit differences only
    r = run(*common, "-s", "1", "-w", "90", "-o", str(a), "--checkpoint", str(ck))   # other image size: refused
    assert r.returncode == 1 and "another image size or seed" in r.stderr
