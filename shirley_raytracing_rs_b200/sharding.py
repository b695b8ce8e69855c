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
    """Weak scaling (bench.py): every rank adds `samples_per_rank` new samples."""
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
