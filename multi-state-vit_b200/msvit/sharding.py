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


def bind_host_thread_to_gpu(device_index: int) -> Optional[List[int]]:
    """Pin the calling host thread to the CPU cores next to GPU `device_index` (the "CPU Affinity" column of
    `nvidia-smi topo -m`), so that pinned buffers allocated afterwards are first-touched on that GPU's NUMA node and
    the H2D / D2H copies do not cross the socket interconnect.  Best effort: returns the core list, or None when the
    topology cannot be read (the caller just keeps its affinity)."""
    import os
    import subprocess
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    cores: List[int] = []
    for line in txt.splitlines():
        f = line.split()
        if not f or f[0] != f"GPU{device_index}":
            continue
        # the affinity field looks like "0-31,64-95"
        for tok in f[1:]:
            if tok and tok[0].isdigit() and all(ch.isdigit() or ch in "-," for ch in tok) and ("-" in tok or "," in tok):
                try:
                    for part in tok.split(","):
                        a, _, b = part.partition("-")
                        cores.extend(range(int(a), int(b or a) + 1))
                    break
                except ValueError:
                    cores = []
        break
    if not cores:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cores) & allowed)
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return use
    except (AttributeError, OSError):
        return None
