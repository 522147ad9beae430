"""Multi-GPU layer of the batched solve: instances are independent, so the batch is
block-partitioned over the ranks (one process per GPU), every rank solves its shard with no
traffic at all, and ONE collective gathers the solutions and the per-instance statistics
afterwards (NCCL all-gather over NVLink; gloo in the CPU tests).  The reference has no
counterpart: it solves one instance per process."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: ceil(batch / world) instances per rank, short (or empty) tail."""
    per = -(-batch // world)
    start = min(batch, rank * per)
    return start, min(batch, start + per) - start


def gather_shards(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gathers equally sized (padded) shards [per, ...] and trims the padding -> [batch, ...]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:batch]
    per = -(-batch // world)
    if local.shape[0] != per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:batch]


def sharded_solve(solve_shard: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                  frames: torch.Tensor, refs: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Solves a global batch: `solve_shard(frames_shard, refs_shard) -> (x_shard, stats_shard)` runs on
    this rank's block of instances; returns the gathered (x [B, N], stats [B, k]) on every rank."""
    batch = frames.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, count = shard_range(batch, rank, world)
    x, stats = solve_shard(frames[start:start + count], refs[start:start + count])
    return gather_shards(x, batch, group), gather_shards(stats, batch, group)
