"""BASELINE configs 4 and 5 on N GPUs of one box (one process per GPU under torchrun):

  C4  scaled random-sphere scene (G=500 -> ~1e6 spheres), 3840x2160, 64 spp, TILE-sharded: rank r renders the
      8x4-pixel tiles t with t % N == r (disjoint pixels; bit-identical to one GPU), frame assembled on rank 0
  C5  Weekend final scene, 3840x2160, 4096 spp in total, SAMPLE-RANGE sharded: rank r renders samples
      [r*4096/N, (r+1)*4096/N) of every pixel; the per-GPU buffers are summed inside the fused peer-memory resolve

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_configs.py [spp_scale] [G]

Time = CUDA events around (render + combine), max over ranks; rays summed over ranks.  Prints one JSON line per config."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import shirley_raytracing_rs_b200 as rt
from shirley_raytracing_rs_b200.sharding import PeerFrame, sample_ranges, tile_shard

F, lib = rt._ffi, rt._ffi.lib
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
G = int(sys.argv[2]) if len(sys.argv) > 2 else 500
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(out_dir, exist_ok=True)


def run(name, scene, cam, total_spp, mode, png):
    W, H = cam.image_width, cam.image_height
    pf = PeerFrame(W, H, local)
    h = scene.device(local)
    if mode == "tiles":
        sc, si = tile_shard(world, rank)
        p = F.RenderParams(samples=total_spp, sample_offset=0, max_depth=50, seed=11, device=-1, shard_count=sc, shard_index=si)
    else:
        sr = sample_ranges(total_spp, world)[rank]
        p = F.RenderParams(samples=sr.samples, sample_offset=sr.sample_offset, max_depth=50, seed=11, device=-1)
    def frame():
        pf.begin_frame()
        F.check(lib.b200rt_render_device(h, C.byref(cam), C.byref(p), pf.accum_ptr, None))
        pf.combine(total_spp)
        st = F.Stats()
        F.check(lib.b200rt_render_device_finish(h, None, C.byref(st)))
        return st
    warm = F.RenderParams(samples=1, max_depth=50, seed=1, device=-1)
    F.check(lib.b200rt_render_device(h, C.byref(cam), C.byref(warm), pf.accum_ptr, None))
    F.check(lib.b200rt_render_device_finish(h, None, None))
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = frame()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), st.kernel_ms], dtype=torch.float64, device=dev)
    r = torch.tensor([float(st.rays), float(st.paths)], dtype=torch.float64, device=dev)
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN); dist.all_reduce(r, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms, rays = float(t[0]), float(r[0])
        info = scene.info(local)
        print(json.dumps({"config": name, "n_gpus": world, "sharding": mode, "image": [W, H], "spp_total": total_spp, "objects": int(info.n_prims),
                          "ms_per_frame": ms, "Mrays_per_s": rays / ms / 1e3, "Msamples_per_s": float(r[1]) / ms / 1e3,
                          "kernel_ms_max": float(t[1]), "kernel_ms_min": float(tmin[1]), "rays": rays}), flush=True)
        rt.write_png(os.path.join(out_dir, png), pf.frame().cpu().numpy())
    pf.close()


camG = rt.camera((0.9 * G, 0.18 * G + 2, 0.35 * G), (0, 0, 0), vfov=30, aperture=0.001, width=3840, aspect_ratio=(16, 9), focus_length=10.0)
run(f"C4 scaled G={G} tile-sharded", rt.Scene.named("scaled", seed=3, param=G), camG, max(1, int(64 * scale)), "tiles", f"c4_n{world}.png")
run("C5 weekend 4K 4096 spp sample-sharded", rt.Scene.named("random", seed=0xDEADBEEF), rt.default_camera(3840, aspect_ratio=(16, 9)),
    max(world, int(4096 * scale)), "samples", f"c5_n{world}.png")
dist.destroy_process_group()
