"""world_size-2 `gloo` test of the multi-rank plumbing bench.py uses at N > 1, on CPU.

There is no GPU here, so each rank renders its shard with the ORACLE (the checker standing
in for the device) — what is under test is the host-side logic: the sample-range plan, the
(seed, pixel, sample) keying that makes ranges composable, and the reduce onto rank 0."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, spp_total, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import shirley_raytracing_rs_b200 as rt
    from shirley_raytracing_rs_b200 import sharding
    from oracle import pyoracle as po
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(48)
    rng = sharding.sample_ranges(spp_total, world)[rank]
    o = po.OracleScene(scene.desc)
    acc, st = o.render(cam, rng.samples, seed=9, sample_offset=rng.sample_offset, threads=1)
    buf = np.zeros((cam.image_height, cam.image_width, 4))
    buf[..., :3] = acc
    buf[..., 3] = rng.samples
    t = torch.from_numpy(buf)
    sharding.reduce_accum(t, dst=0)
    rays = torch.tensor([float(st.rays)], dtype=torch.float64)
    dist.all_reduce(rays)
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # bench.py: time = max over ranks
    dist.barrier()
    if rank == 0:
        np.savez(out_path, accum=t.numpy(), rays=rays.numpy(), ms=ms.numpy())
    dist.destroy_process_group()


def test_sample_range_plan():
    from shirley_raytracing_rs_b200 import sharding
    r = sharding.sample_ranges(500, 8)
    assert sum(x.samples for x in r) == 500 and r[0].sample_offset == 0
    assert all(r[i].sample_offset + r[i].samples == r[i + 1].sample_offset for i in range(7))
    assert [x.samples for x in sharding.sample_ranges(10, 4)] == [3, 3, 2, 2]
    assert sharding.weak_sample_range(500, 3) == sharding.SampleRange(1500, 500)
    assert sharding.tile_shard(4, 2) == (4, 2)
    with pytest.raises(ValueError):
        sharding.tile_shard(2, 2)
    bands = sharding.row_bands(800, 8)                      # fused cross-GPU resolve: one band of rows per rank
    assert bands[0] == (0, 100) and bands[-1] == (700, 800)
    for H, n in ((7, 4), (1, 8), (1080, 3)):
        b = sharding.row_bands(H, n)
        assert b[0][0] == 0 and b[-1][1] == H and all(b[i][1] == b[i + 1][0] for i in range(n - 1))


def test_two_rank_sample_sharding_matches_single_process(tmp_path, rt, po):
    import torch.multiprocessing as mp
    world, spp = 2, 6
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    got = np.load(out)
    scene = rt.Scene.named("random", seed=0xDEADBEEF)
    cam = rt.default_camera(48)
    full, st = po.OracleScene(scene.desc).render(cam, spp, seed=9, threads=1)
    np.testing.assert_allclose(got["accum"][..., :3], full, rtol=1e-12, atol=1e-12)
    assert np.all(got["accum"][..., 3] == spp)
    assert got["rays"][0] == st.rays          # the same paths were traced, just on two ranks
    assert got["ms"][0] == 2.0
