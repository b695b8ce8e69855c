"""N = 2 on real GPUs (run with `gpurun --gpus 2`): sample-range sharding + the fused peer-memory
sum+resolve (b200rt_resolve_peers_rgb8_device over CUDA-IPC peer pointers) against one GPU
rendering all the samples.  Skipped when fewer than two devices are visible."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _strong_worker(rank, world, port, total_spp, width, out_path):
    """ONE frame of `total_spp` samples split into contiguous sample ranges over the ranks (bench.py's strong scaling)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import ctypes as C
    import torch
    import torch.distributed as dist
    import shirley_raytracing_rs_b200 as rt
    from shirley_raytracing_rs_b200.sharding import PeerFrame, sample_ranges
    F = rt._ffi
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(width)
    pf = PeerFrame(cam.image_width, cam.image_height, rank)
    sr = sample_ranges(total_spp, world)[rank]
    rays = 0
    for rep in range(2):
        pf.begin_frame()
        p = F.RenderParams(samples=sr.samples, sample_offset=sr.sample_offset, max_depth=50, seed=31, device=-1)
        F.check(F.lib.b200rt_render_device(scene.device(rank), C.byref(cam), C.byref(p), pf.accum_ptr, None))
        pf.combine(total_spp)
        st = F.Stats()
        F.check(F.lib.b200rt_render_device_finish(scene.device(rank), None, C.byref(st)))
        rays = st.rays
    frame = pf.frame().cpu().numpy().copy() if rank == 0 else None      # frame(): waits for every band, raises on a flag time-out
    tot = torch.tensor([float(rays)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        np.savez(out_path, frame=frame, rays=tot.cpu().numpy())
    pf.close()
    dist.destroy_process_group()


def test_strong_scaled_frame_equals_one_gpu_frame(tmp_path, rt, gpu_required):
    """The BASELINE metric's frame (1200x800, 500 spp) split over two GPUs by sample ranges (250 + 250) is the frame one
    GPU renders alone: the same (pixel, sample) streams, f32 partial sums -> at most one 8-bit level apart, same ray count."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "strong.npz")
    mp.spawn(_strong_worker, args=(2, _free_port(), 500, 1200, out), nprocs=2, join=True)
    got = np.load(out)
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(1200)
    full, st = rt.render(scene, cam, samples=500, seed=31)
    want = rt.resolve_rgb8(full, samples=500)
    d = np.abs(got["frame"].astype(int) - want.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3, (d.max(), (d > 0).mean())
    assert int(got["rays"][0]) == st.rays


def test_bench_line_at_two_gpus(rt, gpu_required):
    """bench.py under torchrun at N = 2 (reduced spp): one JSON line, strong scaling, the in-run correctness checks."""
    import json
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1", "--spp", "20", "--no-other-configs"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["config"]["spp"] == 20 and line["config"]["spp_per_gpu"] == 10
    assert line["combine_check"] == "ok" and line["strong_check"]["max_level_diff"] <= 1
    assert line["e2e"]["value"] > 0 and line["e2e_single_process"]["value"] > 0 and line["gpu_launches"] > 0
    assert "frame_breakdown" in line and line["roofline"]["frac"] > 0


def _worker(rank, world, port, spp, out_path, barrier):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import ctypes as C
    import torch
    import torch.distributed as dist
    import shirley_raytracing_rs_b200 as rt
    from shirley_raytracing_rs_b200.sharding import PeerFrame, weak_sample_range
    F = rt._ffi
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(300)
    W, H = cam.image_width, cam.image_height
    pf = PeerFrame(W, H, rank, barrier=barrier)
    sr = weak_sample_range(spp, rank)
    frames = []
    for rep in range(3):                                   # several frames: buffers and flags are reused
        pf.begin_frame()
        p = F.RenderParams(samples=sr.samples, sample_offset=sr.sample_offset, max_depth=50, seed=4 + rep, device=-1)
        F.check(F.lib.b200rt_render_device(scene.device(rank), C.byref(cam), C.byref(p), pf.accum_ptr, None))
        pf.combine(spp * world)
        st = F.Stats()
        F.check(F.lib.b200rt_render_device_finish(scene.device(rank), None, C.byref(st)))
        torch.cuda.synchronize()
        frames.append(pf.frame().cpu().numpy().copy())
    acc = pf.accum().cpu().numpy().copy()
    assert pf.timed_out() == 0
    gathered = [None] * world
    dist.all_gather_object(gathered, acc)
    if rank == 0:
        np.savez(out_path, frame0=frames[1], frame1=frames[2], **{f"acc{r}": gathered[r] for r in range(world)})
    pf.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("barrier", ["flags", "nccl"])
def test_two_gpu_peer_resolve_matches_one_gpu(tmp_path, rt, gpu_required, barrier):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, spp = 2, 8
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), spp, out, barrier), nprocs=world, join=True)
    got = np.load(out)
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(300)
    # one GPU, all 16 samples of the second frame (seed 6): fixed-point tile sums make the split exact up to f32 adds
    full, _ = rt.render(scene, cam, samples=spp * world, seed=6)
    summed = got["acc0"] + got["acc1"]
    np.testing.assert_allclose(summed[..., :3], full[..., :3], rtol=2e-6, atol=1e-6)
    want = rt.resolve_rgb8(summed.astype(np.float32), samples=spp * world)
    assert np.array_equal(got["frame1"], want)
    assert not np.array_equal(got["frame0"], got["frame1"])      # different seeds, both assembled
    assert got["frame0"].std() > 10


def test_single_process_multi_gpu_render(rt, gpu_required, tmp_path):
    """b200rt_render_rgb8_multi: one process, the samples split over two devices, the sum fused into the resolve over
    peer access — the same picture as one device rendering all the samples (f32 partial sums: +-1 level at most)."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(240)
    one, st1 = rt.render_scene(scene, cam, samples=9, max_reflect=50, output=None, seed=12)
    two, st2 = rt.render_scene_multi(scene, cam, samples=9, max_reflect=50, devices=(0, 1), seed=12)
    assert st2.rays == st1.rays and st2.paths == st1.paths           # the same (pixel, sample) streams, 4 + 5 per device
    d = np.abs(one.astype(int) - two.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    solo, _ = rt.render_scene_multi(scene, cam, samples=9, devices=(1,), seed=12)   # a single, non-default device
    assert np.array_equal(solo, one)
    with pytest.raises(rt.B200rtError):
        rt.render_scene_multi(scene, cam, samples=4, devices=(0, 0))
    cli = os.path.join(ROOT, "ray-cli")
    r = subprocess.run([cli, "-v", "render", "random", "--seed", "5", "-w", "120", "-s", "6", "--gpus", "2", "-o", str(tmp_path / "m.png")], capture_output=True, text=True)
    assert r.returncode == 0 and "on 2 GPUs" in r.stderr, r.stderr
