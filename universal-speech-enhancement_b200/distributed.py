"""Multi-GPU sharding of a batch of clips: one process per GPU, no collective inside the sampling loop, ONE
all-gather of the enhanced waveforms at the end (BASELINE.json north_star).

The reference has no collective on this path at all: under DDP Lightning shards files across ranks and every rank
writes its own wavs (SGMSE_module.py:71-80, loadwav_datamodule.py:53-60).  Works with any torch.distributed backend
(nccl on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n clips: rank r owns [lo, hi); sizes differ by at most one (ragged tails allowed)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_sharded(sample_fn: Callable[[torch.Tensor, int], torch.Tensor], y: torch.Tensor, gather: bool = True,
                   out_shape=None):
    """Run ``sample_fn(y_local, clip0)`` on this rank's shard of ``y`` [B, L] and all-gather the results.

    ``clip0`` is the global index of the shard's first clip: the in-kernel Philox noise is keyed by the global clip
    index, so the gathered result does not depend on the number of ranks (a model-backed ``sample_fn`` should also pass
    ``job_clips=y.shape[0]`` to ``ScoreModel.sample`` so that every shard runs the kernel mode of the whole job).
    ``out_shape``: per-clip shape of
    ``sample_fn``'s result when it differs from ``y.shape[1:]`` (only needed by ranks whose shard is empty, B < world).
    """
    if not (dist.is_available() and dist.is_initialized()):
        return sample_fn(y, 0)
    world, rank = dist.get_world_size(), dist.get_rank()
    B = y.shape[0]
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        out_local = sample_fn(y[lo:hi], lo)
    else:
        # B < world: this rank owns no clip.  Launch nothing (a B = 0 grid is an invalid launch) but still take part in
        # the gather below, otherwise the other ranks would hang in it.
        out_local = y.new_zeros((0,) + tuple(y.shape[1:]), dtype=torch.float32) if out_shape is None else \
            torch.zeros((0,) + tuple(out_shape), dtype=torch.float32, device=y.device)
    if not gather:
        return out_local
    if B == 0:
        return out_local
    if B % world == 0:
        out = torch.empty((B,) + tuple(out_local.shape[1:]), dtype=out_local.dtype, device=out_local.device)
        dist.all_gather_into_tensor(out, out_local.contiguous())
        return out
    # ragged shards: pad to the largest shard, gather, then drop the padding
    mx = (B + world - 1) // world
    pad = torch.zeros((mx,) + tuple(out_local.shape[1:]), dtype=out_local.dtype, device=out_local.device)
    pad[: hi - lo] = out_local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        a, b = shard_range(B, r, world)
        parts.append(bufs[r][: b - a])
    return torch.cat(parts, dim=0)
