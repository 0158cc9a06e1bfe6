"""Batch sharding of the per-image path across ranks.

Images are independent units, so the batch is partitioned contiguously and there is NO collective on the data
path; collectives appear only around it (gathering results for verification, max-over-ranks timing).
Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous partition of `total` units: -> (first, count) of `rank`; sizes differ by at most one and the
    larger shards come first."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_shards(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather row-sharded results (shards made by `shard_bounds`) back into one [total, ...] tensor."""
    world = dist.get_world_size(group)
    counts = [shard_bounds(total, r, world)[1] for r in range(world)]
    width = max(counts)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def max_over_ranks(value: float, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> float:
    """Step time of a sharded job = the slowest rank's time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
