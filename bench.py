#!/usr/bin/env python
"""bench.py — Mrays/s & ms/frame on the Weekend final scene, 1200x800, 500 spp, depth 50
(BASELINE.json metric; workload = configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame: every pixel gets `spp` samples of ray_color(pixel_ray(..)) and the
per-pixel sums land in a float4 accumulation buffer.  A "ray" is one closest-hit query
(one path segment, render.rs:31), counted by the kernel.

  value      whole-job Mrays/s with the scene resident in HBM, timed with CUDA events around
             each step on the launching stream (L2 flushed between steps), max over ranks.
  e2e        the same metric through the reference-facing call sequence with HOST buffers:
             b200rt_scene_create (H2D of the flattened scene) + b200rt_render_rgb8 (render,
             resolve, D2H of the RGB8 frame) per step.
  roofline   FP32-issue roofline of the path-tracing kernel: algorithmic
             lane-ops per ray (24 per box test + 30 per primitive test + 70 fixed, from the
             kernel's own traversal counters) x rays / kernel time, against the FFMA-chain
             ceiling measured live on the same device.  HBM figures are added for
             completeness; the scene (~60 KB) lives in shared memory, so HBM is idle.
  cpu_baseline  the oracle (f64 restatement of the reference, OpenMP over scanlines = the
             reference's rayon row tasks) timed on this box's host cores on a bounded
             sample of the same frame.

Multi-GPU (N > 1, one rank per GPU under torchrun): sample-range sharding — every rank
renders the full frame with its own `spp` samples (sample_offset = rank * spp; streams are
keyed by (pixel, sample) so the union is one N*spp-sample frame).  The per-rank buffers
become one RGB8 frame on rank 0 inside the timed region: by default with the fused
peer-memory kernel (b200rt_resolve_peers_rgb8_device: every rank sums its band of rows from
all ranks' buffers over NVLink P2P, resolves and stores the bytes into rank 0's frame;
`--combine nccl` = one NCCL reduce + resolve on rank 0 instead).  Per-GPU work is fixed:
"scaling": "weak".

`--impl reference` times the reference's own CPU path (the oracle port — the Rust crate
cannot be built in this image) on the same config with all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, SPP, DEPTH, SCENE_SEED = 1200, 500, 50, 0xDEADBEEF
WORKLOAD = "weekend_final_scene_1200x800_500spp_depth50"
METRIC = "Mrays/s, Weekend final scene 1200x800 500spp (ms/frame in ms_per_step)"


# stdout carries exactly ONE JSON line: anything else a library prints there (NCCL's version banner,
# for one) is redirected to stderr.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_ops_per_ray(node_visits, prim_tests, rays):
    """24 lane-ops per box test (each node visit tests both children's boxes),
    30 per primitive test, 70 fixed per segment."""
    return 24.0 * (2.0 * node_visits / rays) + 30.0 * (prim_tests / rays) + 70.0


def algorithmic_bytes_per_ray(node_visits, prim_tests, rays):
    return 64.0 * (node_visits / rays) + 16.0 * (prim_tests / rays) + 32.0


def run_reference(args, rank, world):
    """The reference arm: the oracle port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import shirley_raytracing_rs_b200 as rt
    from oracle import pyoracle as po
    scene = rt.Scene.named("random", seed=SCENE_SEED)
    cam = rt.default_camera(WIDTH)
    o = po.OracleScene(scene.desc, reference_topology=True, precision=64)
    cores = os.cpu_count() or 1
    spp = max(1, args.ref_spp)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        _, st = o.render(cam, spp, max_depth=DEPTH, seed=1000 + i, threads=cores)
        if i >= args.warmup:
            times.append(st.seconds); rays += st.rays
    total = sum(times)
    val = rays / total / 1e6
    sample = f"{WIDTH}x{cam.image_height} full frame at {spp} spp per step (of 500), f64, OpenMP dynamic,1 over scanlines"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scene_seed": SCENE_SEED, "sample": sample, "rays_per_step": rays // max(1, args.steps)},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="samples per pixel per GPU per step (BASELINE: 500)")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--ref-spp", type=int, default=16, help="spp of the bounded CPU sample per step")
    ap.add_argument("--cpu-baseline-spp", type=int, default=128, help="bounded CPU sample: ~10-30 s of host work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--peer-barrier", default="flags", choices=["flags", "nccl"],
                    help="--combine peer: order the ranks with flags in peer memory (default) or a one-element NCCL all_reduce")
    ap.add_argument("--combine", default="peer", choices=["peer", "nccl"],
                    help="N > 1: fused peer-memory reduce+resolve kernel over NVLink (default) or NCCL reduce then resolve on rank 0")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world != 1:
        log(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}")

    import numpy as np
    import torch
    import torch.distributed as dist
    import shirley_raytracing_rs_b200 as rt
    F, lib = rt._ffi, rt._ffi.lib

    if not torch.cuda.is_available() or rt.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    scene = rt.Scene.named("random", seed=SCENE_SEED)
    cam = rt.default_camera(args.width)
    W, H = cam.image_width, cam.image_height
    dscene = scene.device(local_rank)
    info = scene.info(local_rank)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    peer = None
    if world > 1 and args.combine == "peer":
        from shirley_raytracing_rs_b200.sharding import PeerFrame
        peer = PeerFrame(W, H, local_rank, barrier=args.peer_barrier)   # IPC-shared accumulation buffers, flags, the frame on rank 0
        accum = peer.accum()
    else:
        accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    frame_dev = torch.empty((H, W, 3), dtype=torch.uint8, device=dev) if (world > 1 and peer is None and rank == 0) else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def params(step, count=False):
        return F.RenderParams(samples=args.spp, sample_offset=rank * args.spp, max_depth=DEPTH,
                              flags=F.FLAG_COUNT_TRAVERSAL if count else 0, seed=77 + step, device=-1)

    def combine(scene_handle=None):
        """N > 1: the per-rank buffers become ONE RGB8 frame on rank 0."""
        if peer is not None:
            peer.combine(args.spp * world, sptr)          # barrier, fused sum+resolve of this rank's row band into rank 0's frame, barrier
        else:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                F.check(lib.b200rt_resolve_rgb8_device(accum.data_ptr(), W, H, args.spp * world, frame_dev.data_ptr(), sptr))

    def frame(step, count=False):
        """One step with the scene resident in HBM: render (+ at N > 1 the cross-GPU sum and resolve)."""
        p = params(step, count)
        if peer is not None:
            peer.begin_frame(sptr)
        F.check(lib.b200rt_render_device(dscene, C.byref(cam), C.byref(p), accum.data_ptr(), sptr))
        if world > 1:
            combine()

    def finish():
        st = F.Stats()
        F.check(lib.b200rt_render_device_finish(dscene, sptr, C.byref(st)))
        return st

    # ---- counters run (untimed): traversal statistics for the roofline -----------------------
    frame(0, count=True)
    cst = finish()
    ops_per_ray = algorithmic_ops_per_ray(cst.node_visits, cst.prim_tests, cst.rays)
    bytes_per_ray = algorithmic_bytes_per_ray(cst.node_visits, cst.prim_tests, cst.rays)

    # ---- warm-up ---------------------------------------------------------------------------------
    for i in range(args.warmup):
        frame(i); finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    fp32_peak = rt.fp32_peak(local_rank)   # lane-instr/s, measured live (FFMA chain)

    # ---- timed steps -----------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, kernel_ms, rays, launches = [], [], 0, 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                       # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        frame(100 + i)
        e1.record(stream)
        st = finish()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1)); kernel_ms.append(st.kernel_ms)
        rays += st.rays; launches += st.launches
        if world > 1:
            # peer: resolve_peers_kernel + the flag kernels (signal, wait, signal, and the next frame's wait) on every rank;
            # nccl: resolve_kernel on rank 0
            launches += (5 if args.peer_barrier == "flags" else 1) if peer is not None else (1 if rank == 0 else 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))

    # ---- e2e: the reference-facing call with host buffers -----------------------------------------
    rgb = np.empty((H, W, 3), dtype=np.uint8)
    rgb_t = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    e2e_ms, e2e_rays = [], 0
    h2d = int(info.device_bytes)
    d2h = H * W * 3

    def e2e_step(step):
        """scene upload + render + resolve + D2H of the frame, like render_scene (main.rs:65-130)."""
        t0 = time.perf_counter()
        h = C.c_void_p()
        F.check(lib.b200rt_scene_create(scene.desc, local_rank, C.byref(h)))
        try:
            st = F.Stats()
            if world == 1:
                p = params(step)
                F.check(lib.b200rt_render_rgb8(h, C.byref(cam), C.byref(p), rgb_t.data_ptr(), None, C.byref(st)))
            else:
                p = params(step)
                if peer is not None:
                    peer.begin_frame(sptr)
                F.check(lib.b200rt_render_device(h, C.byref(cam), C.byref(p), accum.data_ptr(), sptr))
                combine()
                F.check(lib.b200rt_render_device_finish(h, sptr, C.byref(st)))
                if rank == 0:
                    rgb_t.copy_(peer.frame(sptr) if peer is not None else frame_dev, non_blocking=True)
                torch.cuda.synchronize()
        finally:
            lib.b200rt_scene_destroy(h)
        ms = (time.perf_counter() - t0) * 1e3
        log(f"[rank {rank}] e2e step {step}: {ms:.1f} ms wall (kernel {st.kernel_ms:.1f} ms, device total {st.total_ms:.1f} ms)")
        return ms, st.rays

    e2e_step(0)
    if world > 1:
        dist.barrier()
    for i in range(args.steps):
        ms, r = e2e_step(200 + i)
        e2e_ms.append(ms); e2e_rays += r
    if world > 1:
        dist.barrier()

    # ---- reduce over ranks: max time, sum rays ----------------------------------------------------
    vals = torch.tensor([total_ms, float(sum(e2e_ms)), float(sum(kernel_ms))], dtype=torch.float64, device=dev)
    cnts = torch.tensor([float(rays), float(e2e_rays), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
    total_ms, e2e_total_ms, kernel_total_ms = [float(x) for x in vals.tolist()]
    rays_all, e2e_rays_all, launches_all = [float(x) for x in cnts.tolist()]

    if rank == 0:
        value = rays_all / (total_ms * 1e-3) / 1e6
        e2e_value = e2e_rays_all / (e2e_total_ms * 1e-3) / 1e6
        k_rays_per_s = (rays / len(kernel_ms)) / (np.mean(kernel_ms) * 1e-3)     # this rank's kernel alone
        achieved = k_rays_per_s * ops_per_ray                                      # lane-ops/s
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = {"bound": "fp32-issue", "achieved": achieved / 1e12, "peak": fp32_peak / 1e12, "unit": "Tlane-op/s",
                    "frac": achieved / fp32_peak if fp32_peak else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (ncu --set full,
                    # profiles/r01_ncu_metrics.md): the scene is staged in shared memory, HBM is idle
                    "traffic": 2018560, "traffic_unit": "bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of one 50-spp launch (ncu --set full, profiles/r01_ncu_metrics.md); the 15 MB accumulation buffer is written once and stays in L2",
                    "kernel": "path_trace_kernel_v2", "kernel_ms": float(np.mean(kernel_ms)),
                    "ops_per_ray": ops_per_ray, "node_visits_per_ray": cst.node_visits / cst.rays, "prim_tests_per_ray": cst.prim_tests / cst.rays,
                    "segments_per_sample": cst.rays / cst.paths,
                    "peak_source": "FFMA-chain microbenchmark run live in this process (b200rt_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                    "hbm": {"bound": "hbm", "achieved": k_rays_per_s * bytes_per_ray / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": k_rays_per_s * bytes_per_ray / 1e9 / hbm_peak, "bytes_per_ray": bytes_per_ray,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                            "note": "algorithmic bytes are served from shared memory (scene staged per CTA), not HBM"}}
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD if (args.spp == SPP and args.width == WIDTH) else f"weekend_{W}x{H}_{args.spp}spp_depth{DEPTH}",
                           "scene": "src/scenes.rs random_scene (day), seeded", "scene_seed": SCENE_SEED, "objects": int(info.n_prims),
                           "bvh_nodes": int(info.n_bvh_nodes), "image": [W, H], "spp_per_gpu": args.spp, "max_depth": DEPTH,
                           "parallelism": (f"sample-range x{world}, " + ("fused peer-memory sum+resolve (NVLink P2P)" if peer is not None else "NCCL reduce + resolve on rank 0")) if world > 1 else "single GPU",
                           "l2": "256 MiB buffer written between timed steps (outside the per-step CUDA events)",
                           "rays_per_step": rays_all / args.steps, "Msamples_per_s": (W * H * args.spp * world) / (total_ms / args.steps * 1e-3) / 1e6},
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_total_ms / args.steps,
                        "path": "b200rt_scene_create + b200rt_render_rgb8 (host RGB8 out) per step"},
                "gpu_launches": int(launches_all),
                "clocks": clocks, "roofline": roofline, "wall_s": wall}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as po
            o = po.OracleScene(scene.desc, reference_topology=True, precision=64)
            cores = os.cpu_count() or 1
            o.render(cam, 1, max_depth=DEPTH, seed=5, threads=cores)     # warm
            _, ost = o.render(cam, args.cpu_baseline_spp, max_depth=DEPTH, seed=6, threads=cores)
            _, o1 = o.render(cam, 2, max_depth=DEPTH, seed=7, threads=1)   # the reference's --single-threaded twin (src/main.rs:97-112), 2 spp
            line["cpu_baseline"] = {"value": ost.rays / ost.seconds / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                    "single_thread_value": o1.rays / o1.seconds / 1e6,
                                    "sample": f"{W}x{H} full frame at {args.cpu_baseline_spp} spp (of {args.spp}), f64 oracle, OpenMP dynamic,1 over scanlines, {ost.seconds:.1f} s",
                                    "pops_per_ray": ost.pops / ost.rays, "leaf_tests_per_ray": ost.leaf_tests / ost.rays}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if peer is not None:
        if peer.timed_out():
            log(f"[rank {rank}] WARNING: a peer flag wait timed out (rank {peer.timed_out() - 1} never arrived)")
        peer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
