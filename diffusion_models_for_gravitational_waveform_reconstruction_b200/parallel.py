"""Multi-GPU plumbing of the hot path (one process per GPU, `torch.distributed`).

The reference is single-process (no DDP / NCCL anywhere); what is added here is exactly what the path needs:
  * sampling shards the waveform batch contiguously by rank -- no collective; Philox streams are keyed on the global sample
    index so the result does not depend on the world size;
  * training shards the batch the same way and needs ONE collective per step: all-reduce(sum) of the flat fp32 gradient
    bucket (ParamLayout order, 1 066 952 floats = 4.27 MB for the default model), averaged by the 1/world factor inside the
    fused clip+AdamW+EMA kernel.  The loss is the mean over samples of per-sample masked means (train.py:419-421), so the
    mean of equal-sized rank means is exact.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class ShardInfo:
    rank: int
    world: int
    start: int          # global index of this rank's first sample (Philox `sample0`)
    count: int


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, start+count) of `n_total` samples; the first n_total % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    q, r = divmod(int(n_total), int(world))
    start = rank * q + min(rank, r)
    return start, q + (1 if rank < r else 0)


def current_shard(n_total: int, group=None) -> ShardInfo:
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    s, c = shard_range(n_total, rank, world)
    return ShardInfo(rank, world, s, c)


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the flat gradient bucket over the ranks (NCCL on GPUs, gloo in the CPU tests).  No-op for world 1."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def grad_scale(group=None) -> float:
    """Factor applied to the summed bucket by gw_adamw_ema (hyper[6])."""
    if dist.is_available() and dist.is_initialized():
        return 1.0 / dist.get_world_size(group)
    return 1.0
