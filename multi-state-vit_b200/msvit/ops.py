"""torch custom operators over the C ABI (`msvit::*`): thin, allocation in torch, compute in libmsvit.so.

Every operator is CUDA-only.  Inputs must be contiguous CUDA tensors; the kernels run on the current
stream of the tensor's device.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def lda_of(n: int) -> int:
    return (n + 3) & ~3


def affinity_stride(N: int) -> int:
    """Per-image capacity (floats) of the segment-packed affinity buffer."""
    return (N * (N + 3) + 3) & ~3


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"msvit kernels take float32 or bfloat16 tokens, got {t.dtype}")


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("msvit operators are CUDA (sm_100a) only; there is no CPU fallback")
        if not t.is_contiguous():
            raise ValueError("msvit operators need contiguous tensors")


@torch.library.custom_op("msvit::affinity_degree", mutates_args=(), device_types="cuda")
def affinity_degree(x: torch.Tensor, S: int, N: int, mode: int, gamma: float, scale: float,
                    seg_off: Optional[torch.Tensor], a_off: Optional[torch.Tensor], a_numel: int,
                    want_affinity: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [rows, D] -> (A flat [a_numel] (segment blocks, see msvit.h), deg [rows])."""
    _need_cuda(x, seg_off, a_off)
    rows, D = x.shape
    with torch.cuda.device(x.device):
        A = torch.empty(a_numel if want_affinity else 0, dtype=torch.float32, device=x.device)
        deg = torch.empty(rows, dtype=torch.float32, device=x.device)
        code = _lib.load().msvit_affinity_degree(_ptr(x), _dtype_code(x), _ptr(A) if want_affinity else None, _ptr(deg),
                                                 rows, S, N, D, mode, gamma, scale, _ptr(seg_off), _ptr(a_off),
                                                 _stream(x))
    _lib.check(code, "msvit_affinity_degree")
    return A, deg


@affinity_degree.register_fake
def _(x, S, N, mode, gamma, scale, seg_off, a_off, a_numel, want_affinity):
    return x.new_empty(a_numel if want_affinity else 0, dtype=torch.float32), x.new_empty(x.shape[0], dtype=torch.float32)


@torch.library.custom_op("msvit::ncut_eig", mutates_args=(), device_types="cuda")
def ncut_eig(A: torch.Tensor, deg: torch.Tensor, S: int, N: int, k: int, block: int, max_iter: int, tol: float,
             lam_floor: float, n_converge: int, seg_off: Optional[torch.Tensor],
             a_off: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (V [rows, k], lam [S, k], iters [S] int32)."""
    _need_cuda(A, deg, seg_off, a_off)
    rows = deg.shape[0]
    with torch.cuda.device(A.device):
        V = torch.empty(rows, k, dtype=torch.float32, device=A.device)
        lam = torch.empty(S, k, dtype=torch.float32, device=A.device)
        iters = torch.empty(S, dtype=torch.int32, device=A.device)
        code = _lib.load().msvit_ncut_eig(_ptr(A), _ptr(deg), _ptr(V), _ptr(lam), _ptr(iters), rows, S, N, k, block,
                                          max_iter, tol, lam_floor, n_converge, _ptr(seg_off), _ptr(a_off),
                                          _stream(A))
    _lib.check(code, "msvit_ncut_eig")
    return V, lam, iters


@ncut_eig.register_fake
def _(A, deg, S, N, k, block, max_iter, tol, lam_floor, n_converge, seg_off, a_off):
    rows = deg.shape[0]
    return (A.new_empty(rows, k), A.new_empty(S, k), A.new_empty(S, dtype=torch.int32))


@torch.library.custom_op("msvit::kmeans", mutates_args=(), device_types="cuda")
def kmeans(V: torch.Tensor, lam: Optional[torch.Tensor], weight: Optional[torch.Tensor], init: Optional[torch.Tensor],
           S: int, N: int, n_clusters: int, eig_threshold: float, max_iter: int,
           seg_off: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (labels [rows] int32 local canonical ids, n_child [S] int32, centres [S, Kmax, Kmax])."""
    _need_cuda(V, lam, weight, init, seg_off)
    rows, ldv = V.shape
    Kmax = n_clusters if n_clusters > 0 else min(ldv, _lib.MAX_EIG_BLOCK)
    with torch.cuda.device(V.device):
        labels = torch.empty(rows, dtype=torch.int32, device=V.device)
        n_child = torch.empty(S, dtype=torch.int32, device=V.device)
        centres = torch.empty(S, Kmax, Kmax, dtype=torch.float32, device=V.device)
        code = _lib.load().msvit_kmeans(_ptr(V), _ptr(lam), _ptr(weight), _ptr(init), _ptr(labels), _ptr(n_child),
                                        _ptr(centres), rows, S, N, ldv, n_clusters, eig_threshold, max_iter,
                                        _ptr(seg_off), _stream(V))
    _lib.check(code, "msvit_kmeans")
    return labels, n_child, centres


@kmeans.register_fake
def _(V, lam, weight, init, S, N, n_clusters, eig_threshold, max_iter, seg_off):
    rows, ldv = V.shape
    Kmax = n_clusters if n_clusters > 0 else min(ldv, _lib.MAX_EIG_BLOCK)
    return (V.new_empty(rows, dtype=torch.int32), V.new_empty(S, dtype=torch.int32), V.new_empty(S, Kmax, Kmax))


@torch.library.custom_op("msvit::pool", mutates_args=(), device_types="cuda")
def pool(x: torch.Tensor, labels: torch.Tensor, K: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [B, N, D], labels [B, N] int64 -> (pooled [B, K, D] fp32, counts [B, K] int32)."""
    _need_cuda(x, labels)
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    B, N, D = x.shape
    with torch.cuda.device(x.device):
        pooled = torch.empty(B, K, D, dtype=torch.float32, device=x.device)
        counts = torch.empty(B, K, dtype=torch.int32, device=x.device)
        code = _lib.load().msvit_pool(_ptr(x), _dtype_code(x), _ptr(labels), _ptr(pooled), _ptr(counts), B, N, D, K,
                                      _stream(x))
    _lib.check(code, "msvit_pool")
    return pooled, counts


@pool.register_fake
def _(x, labels, K):
    B, N, D = x.shape
    return x.new_empty(B, K, D, dtype=torch.float32), x.new_empty(B, K, dtype=torch.int32)


@torch.library.custom_op("msvit::build_segments", mutates_args=(), device_types="cuda")
def build_segments(parent_indices: torch.Tensor, P: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """parent_indices [B, N] int64 in [0, P) -> (perm [B*N] int32, seg_off [B*P+1] int32, a_off [B*P+1] int64)."""
    _need_cuda(parent_indices)
    if parent_indices.dtype != torch.int64:
        raise TypeError("parent_indices must be int64")
    B, N = parent_indices.shape
    dev = parent_indices.device
    with torch.cuda.device(dev):
        perm = torch.empty(B * N, dtype=torch.int32, device=dev)
        seg_off = torch.empty(B * P + 1, dtype=torch.int32, device=dev)
        a_off = torch.empty(B * P + 1, dtype=torch.int64, device=dev)
        code = _lib.load().msvit_build_segments(_ptr(parent_indices), _ptr(perm), _ptr(seg_off), _ptr(a_off), B, N, P,
                                                _stream(parent_indices))
    _lib.check(code, "msvit_build_segments")
    return perm, seg_off, a_off


@build_segments.register_fake
def _(parent_indices, P):
    B, N = parent_indices.shape
    return (parent_indices.new_empty(B * N, dtype=torch.int32), parent_indices.new_empty(B * P + 1, dtype=torch.int32),
            parent_indices.new_empty(B * P + 1, dtype=torch.int64))


@torch.library.custom_op("msvit::gather_rows", mutates_args=(), device_types="cuda")
def gather_rows(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """xs[j] = x[perm[j]] for a [rows, D] matrix."""
    _need_cuda(x, perm)
    rows, D = x.shape
    with torch.cuda.device(x.device):
        xs = torch.empty_like(x)
        code = _lib.load().msvit_gather_rows(_ptr(x), _dtype_code(x), _ptr(perm), _ptr(xs), rows, D, _stream(x))
    _lib.check(code, "msvit_gather_rows")
    return xs


@gather_rows.register_fake
def _(x, perm):
    return torch.empty_like(x)


@torch.library.custom_op("msvit::compose_labels", mutates_args=(), device_types="cuda")
def compose_labels(labels_sorted: torch.Tensor, n_child: torch.Tensor, perm: Optional[torch.Tensor],
                   seg_off: Optional[torch.Tensor], B: int, N: int, P: int) -> torch.Tensor:
    """-> child_indices [B, N] int64 (contiguous per image, ranges ordered by parent)."""
    _need_cuda(labels_sorted, n_child, perm, seg_off)
    dev = labels_sorted.device
    with torch.cuda.device(dev):
        child = torch.empty(B, N, dtype=torch.int64, device=dev)
        code = _lib.load().msvit_compose_labels(_ptr(labels_sorted), _ptr(n_child), _ptr(perm), _ptr(seg_off),
                                                _ptr(child), B, N, P, _stream(labels_sorted))
    _lib.check(code, "msvit_compose_labels")
    return child


@compose_labels.register_fake
def _(labels_sorted, n_child, perm, seg_off, B, N, P):
    return labels_sorted.new_empty(B, N, dtype=torch.int64)
