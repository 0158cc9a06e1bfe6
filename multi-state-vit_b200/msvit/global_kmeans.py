"""Dataset-level ("DeepCluster-style") k-means with the rows sharded over the GPUs of a box.

Reference: the flattened-batch clustering of model/clustering/modeling_spectral.py:254-256 (all B*N tokens at once)
and its KMeans(n_clusters).fit_predict call sites (:90, :130-133); BASELINE.json configs[4] (1M x 768 features,
k = 1000, 2/4/8 GPUs).

Each rank owns a contiguous row shard; centroids are replicated.  One Lloyd iteration:

    assign (tcgen05 contraction + fused argmin)  ->  stable counting sort of the row ids by label
    ->  per-centroid sums of the member rows (fixed order, no atomics)  ->  packed [k, D+1] = sums | counts
    ->  ONE all-reduce(sum) of the packed buffer over the process group (NCCL over NVLink; 3.08 MB at k=1000, D=768)
    ->  identical division on every rank

This is the only exchange step of the whole repository; the per-image path has none.  `lloyd` is the host loop,
written against two callables so that the world-size-2 gloo test can drive it on CPU with the oracle's arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _lib, ops


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def lloyd(local_step: Callable[[], torch.Tensor], finalize: Callable[[torch.Tensor], None], iters: int,
          group: Optional["dist.ProcessGroup"] = None) -> None:
    """Host loop of the sharded Lloyd iteration.

    local_step() -> packed [k, D+1] (this rank's centroid sums | counts for the CURRENT centroids);
    finalize(packed) consumes the globally reduced buffer and installs the next centroids.
    The all-reduce is skipped for a single rank."""
    world = _world(group)
    for _ in range(int(iters)):
        packed = local_step()
        if world > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        finalize(packed)


def broadcast_init(local_rows: torch.Tensor, k: int, group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """Initial centres = the first k rows of the GLOBAL matrix (rank 0's shard), replicated on every rank."""
    world = _world(group)
    D = local_rows.shape[1]
    init = torch.empty(k, D, dtype=torch.float32, device=local_rows.device)
    rank = dist.get_rank(group) if world > 1 else 0
    if rank == 0:
        if local_rows.shape[0] < k:
            raise ValueError(f"rank 0 holds {local_rows.shape[0]} rows, fewer than k={k} initial centres")
        init.copy_(local_rows[:k].float())
    if world > 1:
        dist.broadcast(init, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return init


@dataclass
class GlobalKMeansResult:
    centroids: torch.Tensor   # [k, D] fp32, identical on every rank
    labels: torch.Tensor      # [n_local] int64: assignment of this rank's rows to the centroids BEFORE the last update
    counts: torch.Tensor      # [k] int64 global member counts of that assignment


class GlobalKMeansPlan:
    """Pre-allocated buffers of the sharded Lloyd iteration for a fixed shard shape (n rows, D columns, k centres)."""

    def __init__(self, n: int, D: int, k: int, dtype: torch.dtype = torch.bfloat16, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("msvit.global_kmeans runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"features must be float32 or bfloat16, got {dtype}")
        self.lib = _lib.load()
        self.n, self.D, self.k, self.dtype, self.device = int(n), int(D), int(k), dtype, dev
        self.code = _lib.F32 if dtype == torch.float32 else _lib.BF16
        with torch.cuda.device(dev):
            self.labels = torch.empty(max(self.n, 1), dtype=torch.int32, device=dev)
            self.perm = torch.empty(max(self.n, 1), dtype=torch.int32, device=dev)
            self.seg_off = torch.empty(self.k + 1, dtype=torch.int32, device=dev)
            self.acc_ws_bytes = int(self.lib.msvit_gkm_accumulate_workspace_bytes(self.k, self.D))
            self.ws_bytes = max(int(self.lib.msvit_gkm_workspace_bytes(self.n, self.k)), self.acc_ws_bytes)
            self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)
            self.packed = torch.empty(self.k, self.D + 1, dtype=torch.float32, device=dev)
            self.centroids = torch.empty(self.k, self.D, dtype=torch.float32, device=dev)
            self.centroids_op = self.centroids if dtype == torch.float32 else torch.empty(self.k, self.D, dtype=dtype,
                                                                                         device=dev)
        self.allreduce_bytes = self.packed.numel() * 4

    def set_centroids(self, c: torch.Tensor) -> None:
        self.centroids.copy_(c.to(torch.float32))
        if self.centroids_op is not self.centroids:
            self.centroids_op.copy_(self.centroids)

    def local_step(self, x: torch.Tensor, events=None) -> torch.Tensor:
        """assign -> sort -> accumulate for this rank's rows against the current centroids; returns `packed`."""
        if tuple(x.shape) != (self.n, self.D) or x.dtype != self.dtype or x.device != self.device:
            raise ValueError(f"plan was built for {(self.n, self.D)} {self.dtype} on {self.device}")
        if not x.is_contiguous():
            raise ValueError("features must be contiguous")
        lib, check, p = self.lib, _lib.check, ops._ptr
        st = torch.cuda.current_stream(self.device).cuda_stream

        def mark(i):
            if events is not None:
                events[i].record()

        with torch.cuda.device(self.device):
            mark(0)
            check(lib.msvit_gkm_assign(p(x), self.code, p(self.centroids_op), p(self.labels), None, self.n, self.k,
                                       self.D, st), "msvit_gkm_assign")
            mark(1)
            check(lib.msvit_gkm_sort(p(self.labels), self.n, self.k, p(self.perm), p(self.seg_off), p(self.ws),
                                     self.ws_bytes, st), "msvit_gkm_sort")
            mark(2)
            check(lib.msvit_gkm_accumulate(p(x), self.code, p(self.perm), p(self.seg_off), p(self.packed), self.n,
                                           self.k, self.D, p(self.ws), self.ws_bytes, st), "msvit_gkm_accumulate")
            mark(3)
        return self.packed

    def finalize(self, packed: torch.Tensor) -> None:
        lib, check, p = self.lib, _lib.check, ops._ptr
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            op = None if self.centroids_op is self.centroids else p(self.centroids_op)
            check(lib.msvit_gkm_finalize(p(packed), p(self.centroids), op, self.code, self.k, self.D, st),
                  "msvit_gkm_finalize")


def global_kmeans(features: torch.Tensor, k: int, iters: int, process_group: Optional["dist.ProcessGroup"] = None,
                  init: Optional[torch.Tensor] = None) -> GlobalKMeansResult:
    """Lloyd k-means over row-sharded `features` [n_local, D] (float32 or bfloat16, CUDA).

    init: [k, D] initial centres (identical on every rank); default = the first k rows of the global matrix.
    Returns centroids after `iters` updates, the labels of the last assignment and its global counts."""
    if features.dim() != 2:
        raise ValueError("features must be [rows, D]")
    if not features.is_cuda:
        raise RuntimeError("msvit.global_kmeans runs on CUDA (sm_100a) only; there is no CPU fallback")
    features = features.contiguous()
    n, D = features.shape
    plan = GlobalKMeansPlan(n, D, k, features.dtype, features.device)
    plan.set_centroids(init if init is not None else broadcast_init(features, k, process_group))
    counts = torch.zeros(k, dtype=torch.int64, device=features.device)

    def finalize(packed):
        counts.copy_(packed[:, D].round().to(torch.int64))
        plan.finalize(packed)

    lloyd(lambda: plan.local_step(features), finalize, iters, process_group)
    return GlobalKMeansResult(plan.centroids, plan.labels[:n].long(), counts)
