"""Multi-GPU sharding plans (one process per GPU, torch.distributed for the plumbing).

The reference's only parallelism is rayon over scanlines (src/main.rs:118-125); pixels and
samples are independent (render.rs:58-69), and this backend keys every random stream by
(seed, pixel, sample index).  So a frame shards two ways without touching the kernel:

* sample-range sharding — rank r renders samples [offset_r, offset_r + n_r) of EVERY pixel
  into its own float4 accumulation buffer; the buffers are summed (one NCCL reduce);
* interleaved tile sharding — rank r renders the 8x4-pixel tiles t with t % world == r;
  ranks own disjoint pixels, so the sum of the buffers is bit-identical to one GPU.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class SampleRange:
    sample_offset: int
    samples: int


def sample_ranges(total_samples: int, world: int) -> list[SampleRange]:
    """Split `total_samples` into `world` contiguous ranges (first ranks take the remainder)."""
    if world < 1 or total_samples < 0:
        raise ValueError("world >= 1 and total_samples >= 0 required")
    base, rem = divmod(total_samples, world)
    out, off = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append(SampleRange(off, n))
        off += n
    return out


def weak_sample_range(samples_per_rank: int, rank: int) -> SampleRange:
    """Weak scaling (bench.py --scaling weak): every rank adds `samples_per_rank` new samples."""
    return SampleRange(rank * samples_per_rank, samples_per_rank)


def tile_shard(world: int, rank: int) -> tuple[int, int]:
    """(shard_count, shard_index) for B200rtRenderParams: interleaved 8x4 tiles."""
    if not 0 <= rank < world:
        raise ValueError("0 <= rank < world required")
    return world, rank


def reduce_accum(accum, dst: int = 0):
    """Sum per-rank accumulation buffers onto `dst` (torch tensor, any backend)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def row_bands(height: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row bands [begin, end), one per rank, for the fused cross-GPU resolve."""
    if world < 1 or height < 0:
        raise ValueError("world >= 1 and height >= 0 required")
    return [(r * height // world, (r + 1) * height // world) for r in range(world)]


class _DevicePtr:
    """A raw device allocation seen through __cuda_array_interface__ (torch.as_tensor accepts it)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerFrame:
    """Per-rank accumulation buffers visible to every rank of one node (CUDA IPC over NVLink)
    plus the RGB8 frame on `root`, for b200rt_resolve_peers_rgb8_device: the cross-GPU sum is
    fused into the resolve and every rank assembles its band of rows directly into the root's
    frame (include/b200rt.h).  One process per GPU; handles travel over torch.distributed once.
    Per-frame ordering between the ranks is a flag barrier over the same peer memory
    (b200rt_peer_signal_device / _wait_device): no library collective on the data path
    (barrier="nccl" uses a one-element all_reduce instead, for comparison).

        pf = PeerFrame(W, H, device_index)
        per frame:  pf.begin_frame(stream)            # peers have finished reading last frame's buffers
                    render into pf.accum_ptr on `stream`
                    pf.combine(total_samples, stream) # collective: sum + resolve of this rank's band into root's frame
        frame = pf.frame(stream)                      # on root: (H, W, 3) uint8 torch tensor on the device
    """

    def __init__(self, width: int, height: int, device: int, root: int = 0, group=None, barrier: str = "flags"):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _ffi as F
        self._F, self._C, self._torch, self._dist = F, C, torch, dist
        self.W, self.H, self.device, self.root, self.group = width, height, device, root, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("at most 16 peers")
        if barrier not in ("flags", "nccl"):
            raise ValueError("barrier must be 'flags' or 'nccl'")
        self.barrier = barrier
        lib = F.lib
        self._own = []
        def create(nbytes):
            p, h = C.c_void_p(), (C.c_uint8 * 64)()
            F.check(lib.b200rt_peer_buffer_create(device, nbytes, C.byref(p), h))
            self._own.append(p)
            return p, bytes(h)
        acc_p, acc_h = create(width * height * 16)
        flag_p, flag_h = create(256)
        rgb_p, rgb_h = create(width * height * 3) if self.rank == root else (None, None)
        handles = [None] * self.world
        dist.all_gather_object(handles, (acc_h, rgb_h, flag_h), group=group)
        self._opened = []
        def open_(h):
            p, buf = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(h)
            F.check(lib.b200rt_peer_buffer_open(device, buf, C.byref(p)))
            self._opened.append(p)
            return p
        self.accum_ptrs = [acc_p if r == self.rank else open_(handles[r][0]) for r in range(self.world)]
        self.flag_ptrs = [flag_p if r == self.rank else open_(handles[r][2]) for r in range(self.world)]
        self.frame_ptr = rgb_p if self.rank == root else open_(handles[root][1])
        self.accum_ptr, self.flag_ptr = acc_p, flag_p
        self._ptr_array = (C.c_void_p * self.world)(*[p.value for p in self.accum_ptrs])
        self._flag_array = (C.c_void_p * self.world)(*[p.value for p in self.flag_ptrs])
        self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device))
        self.band = row_bands(height, self.world)[self.rank]
        self.epoch = 0
        self.timeout_ms = 10000
        dist.barrier(group=group)          # every rank has opened every handle before any kernel touches a peer

    # -- ordering ------------------------------------------------------------------------------------------
    def _signal(self, slot, stream_ptr):
        F = self._F
        F.check(F.lib.b200rt_peer_signal_device(self._flag_array, self.world, self.rank, slot, self.epoch, stream_ptr))

    def _wait(self, slot, stream_ptr):
        F = self._F
        F.check(F.lib.b200rt_peer_wait_device(self.flag_ptr, self.world, slot, self.epoch, self.timeout_ms, stream_ptr))

    def sync(self):
        """Stream-ordered barrier through NCCL (barrier='nccl'): kernels enqueued after it start after every rank's earlier work."""
        self._dist.all_reduce(self._flag, group=self.group)

    def begin_frame(self, stream_ptr=None):
        """Before rendering into accum_ptr again: every peer has finished reading the previous frame's buffers."""
        if self.barrier == "flags" and self.epoch > 0:
            self._wait(1, stream_ptr)

    def combine(self, total_samples: int, stream_ptr=None, events=None):
        """Collective.  Orders this rank's render before the peers' reads, resolves this rank's band of rows from
        all ranks' buffers into the root's frame, and publishes 'done reading'.  `events(k)`, if given, is called
        after stage k (0: the ranks are ordered, 1: the band is resolved, 2: 'done' is published) — bench.py's
        per-frame cost breakdown records a CUDA event there."""
        F = self._F
        self.epoch += 1
        if self.barrier == "flags":
            self._signal(0, stream_ptr)
            self._wait(0, stream_ptr)
        else:
            self.sync()
        if events:
            events(0)
        # the flag array rides along: after a wait that timed out the kernel stores nothing (no frame from incomplete buffers)
        F.check(F.lib.b200rt_resolve_peers_rgb8_device(self._ptr_array, self.world, self.W, self.H, total_samples,
                                                      self.band[0], self.band[1], self.frame_ptr,
                                                      self.flag_ptr if self.barrier == "flags" else None, stream_ptr))
        if events:
            events(1)
        if self.barrier == "flags":
            self._signal(1, stream_ptr)
        else:
            self.sync()
        if events:
            events(2)

    def accum(self):
        """This rank's accumulation buffer as an (H, W, 4) float32 torch tensor (no copy)."""
        return self._torch.as_tensor(_DevicePtr(self.accum_ptr.value, (self.H, self.W, 4), "<f4"), device=self._torch.device("cuda", self.device))

    def wait_frame(self, stream_ptr=None):
        """Stream-ordered: work enqueued after this starts after EVERY rank's band has landed in the root's frame."""
        if self.barrier == "flags" and self.epoch > 0:
            self._wait(1, stream_ptr)

    def frame_tensor(self):
        """The frame buffer as an (H, W, 3) uint8 torch tensor, no ordering and no check (pair with wait_frame + check)."""
        return self._torch.as_tensor(_DevicePtr(self.frame_ptr.value, (self.H, self.W, 3), "|u1"), device=self._torch.device("cuda", self.device))

    def frame(self, stream_ptr=None):
        """The assembled RGB8 frame, (H, W, 3) uint8, top row first; ordered after every rank's band (call after combine).
        Raises if a flag wait of this or an earlier frame gave up on a peer (the frame would be incomplete)."""
        self.wait_frame(stream_ptr)
        self.check()
        return self._torch.as_tensor(_DevicePtr(self.frame_ptr.value, (self.H, self.W, 3), "|u1"), device=self._torch.device("cuda", self.device))

    def timed_out(self) -> int:
        """0, or 1 + the rank a flag wait gave up on (a peer died or never called combine).  Synchronises the device;
        a reported time-out is cleared, so the next frame starts clean."""
        out = self._C.c_uint32()
        self._F.check(self._F.lib.b200rt_peer_timed_out(self.flag_ptr, self._C.byref(out)))
        return out.value

    def check(self) -> None:
        """Raise if a flag wait timed out since the last check: frames combined in between are not valid."""
        t = self.timed_out() if self.barrier == "flags" else 0
        if t:
            raise RuntimeError(f"rank {self.rank}: peer flag wait timed out after {self.timeout_ms} ms (rank {t - 1} never arrived); "
                               "the frame was not assembled")

    def close(self):
        F = self._F
        self._torch.cuda.synchronize()
        self._dist.barrier(group=self.group)          # nobody may still be reading a buffer about to be freed
        for p in self._opened:
            F.check(F.lib.b200rt_peer_buffer_close(self.device, p))
        self._dist.barrier(group=self.group)
        for p in self._own:
            F.check(F.lib.b200rt_peer_buffer_destroy(self.device, p))
        self._opened, self._own = [], []
