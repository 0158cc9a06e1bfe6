"""Functional surface of the token-grouping path (added next to the reference's module interface).

    cluster_tokens(x, parent_indices=None, ...) -> ClusterOutput
    pool(x, labels, K)                          -> (pooled, counts)

`cluster_tokens` is the whole hot path for one batch shard: affinity + degree (tcgen05), top-k NCut
eigenvectors, k-means on the embedding, label composition, cluster-mean pooling.  It enqueues a fixed
sequence of kernels on the current stream and never synchronises with the host, except for the single
`.max()` read the reference also does (modeling_spectral.py:80) when the number of parents is not given.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib, ops


@dataclass
class ClusterOutput:
    labels: torch.Tensor            # [B, N] int64 child cluster ids (contiguous per image, ordered by parent)
    pooled: Optional[torch.Tensor]  # [B, K_pool, D] fp32 cluster means ("multi-state tokens")
    counts: Optional[torch.Tensor]  # [B, K_pool] int32
    eigvecs: torch.Tensor           # [B, N, k] fp32, row i = embedding of token i inside its segment
    eigvals: torch.Tensor           # [B, P, k] fp32 (P parents per image)
    n_child: torch.Tensor           # [B, P] int32 children per parent
    degree: torch.Tensor            # [B, N] fp32 NCut degree
    iters: torch.Tensor             # [B, P] int32 eigensolver iterations
    affinity: Optional[torch.Tensor] = None  # [B, N, N] fp32 (single-parent case with N % 4 == 0 only)


def default_block(k: int, oversample: int = 8) -> int:
    """Subspace width: k wanted + oversampling, multiple of 4, at most MAX_EIG_BLOCK."""
    m = (k + oversample + 3) & ~3
    m = min(m, _lib.MAX_EIG_BLOCK)
    if m < k:
        raise ValueError(f"ncut_dim={k} exceeds the supported subspace width {_lib.MAX_EIG_BLOCK}")
    return max(m, (k + 3) & ~3)


def cluster_tokens(x: torch.Tensor, parent_indices: Optional[torch.Tensor] = None, *, ncut_dim: int,
                   n_clusters: Optional[int] = None, eigenvalue_threshold: Optional[float] = None,
                   mode: str = "rbf", gamma: float = 3.0, scale: Optional[float] = None,
                   n_parents: Optional[int] = None, kmeans_iters: int = 100, eig_iters: int = 60,
                   eig_tol: float = 2e-5, oversample: int = 8, pool_k: Optional[int] = None,
                   want_pool: bool = True, keep_affinity: bool = False) -> ClusterOutput:
    if x.dim() != 3:
        raise ValueError("x must be [batch, tokens, hidden]")
    if mode not in _lib.DIST:
        raise ValueError(f"unknown distance {mode!r}")
    if n_clusters is None and eigenvalue_threshold is None:
        raise ValueError("give n_clusters or eigenvalue_threshold")
    if not x.is_cuda:
        raise RuntimeError("msvit.cluster_tokens runs on CUDA (sm_100a) only; there is no CPU fallback")
    B, N, D = x.shape
    x = x.contiguous()
    k = int(ncut_dim)
    block = default_block(k, oversample)
    s = float(D) if scale is None else float(scale)
    flat = x.view(B * N, D)

    if parent_indices is None:
        P = 1
    else:
        if parent_indices.shape != (B, N):
            raise ValueError("parent_indices must be [batch, tokens]")
        parent_indices = parent_indices.contiguous()
        # the reference reads this on the host too (modeling_spectral.py:80); pass n_parents to skip the sync
        P = int(n_parents) if n_parents is not None else int(parent_indices.max().item()) + 1

    if P == 1:
        perm = seg_off = a_off = None
        xs = flat
        S = B
        a_numel = B * N * ops.lda_of(N)
    else:
        perm, seg_off, a_off = ops.build_segments(parent_indices, P)
        xs = ops.gather_rows(flat, perm)
        S = B * P
        a_numel = B * ops.affinity_stride(N)

    A, deg = ops.affinity_degree(xs, S, N, _lib.DIST[mode], float(gamma), s, seg_off, a_off, a_numel, True)
    V, lam, iters = ops.ncut_eig(A, deg, S, N, k, block, int(eig_iters), float(eig_tol), seg_off, a_off)
    nk = int(n_clusters) if n_clusters is not None else 0
    thr = float(eigenvalue_threshold) if eigenvalue_threshold is not None else 0.0
    labels_sorted, n_child, _ = ops.kmeans(V, lam, deg, None, S, N, nk, thr, int(kmeans_iters), seg_off)
    child = ops.compose_labels(labels_sorted, n_child, perm, seg_off, B, N, P)

    if perm is not None:
        idx = perm.long()
        V_tok = torch.empty_like(V)
        V_tok[idx] = V
        deg_tok = torch.empty_like(deg)
        deg_tok[idx] = deg
    else:
        V_tok, deg_tok = V, deg

    pooled = counts = None
    if want_pool:
        Kp = pool_k if pool_k is not None else P * (nk if nk > 0 else k)
        pooled, counts = ops.pool(x, child, int(Kp))

    aff = None
    if keep_affinity and P == 1 and N % 4 == 0:
        aff = A.view(B, N, N)
    return ClusterOutput(labels=child, pooled=pooled, counts=counts, eigvecs=V_tok.view(B, N, k),
                         eigvals=lam.view(B, P, k), n_child=n_child.view(B, P), degree=deg_tok.view(B, N),
                         iters=iters.view(B, P), affinity=aff)


def affinity(x: torch.Tensor, mode: str = "rbf", gamma: float = 3.0, scale: Optional[float] = None):
    """x [B, N, D] -> (A [B, N, lda] fp32 (lda = N rounded up to 4, pad columns are 0), deg [B, N])."""
    B, N, D = x.shape
    x = x.contiguous()
    s = float(D) if scale is None else float(scale)
    lda = ops.lda_of(N)
    A, deg = ops.affinity_degree(x.view(B * N, D), B, N, _lib.DIST[mode], float(gamma), s, None, None, B * N * lda, True)
    return A.view(B, N, lda), deg.view(B, N)


def ncut_eig(A: torch.Tensor, deg: torch.Tensor, k: int, *, max_iter: int = 60, tol: float = 2e-5, oversample: int = 8):
    """A [B, N, lda] (as returned by `affinity`), deg [B, N] -> (V [B, N, k], lam [B, k], iters [B])."""
    B, N, lda = A.shape
    if lda != ops.lda_of(N):
        raise ValueError("A must be [B, N, (N+3)&~3]")
    V, lam, iters = ops.ncut_eig(A.contiguous().view(-1), deg.contiguous().view(-1), B, N, int(k),
                                 default_block(int(k), oversample), int(max_iter), float(tol), None, None)
    return V.view(B, N, k), lam, iters


def kmeans(V: torch.Tensor, n_clusters: Optional[int] = None, *, eigvals: Optional[torch.Tensor] = None,
           eigenvalue_threshold: float = 0.0, weight: Optional[torch.Tensor] = None,
           init: Optional[torch.Tensor] = None, max_iter: int = 100):
    """V [B, N, k] -> (labels [B, N] int64, n_child [B] int32, centres [B, K, K])."""
    B, N, k = V.shape
    nk = int(n_clusters) if n_clusters is not None else 0
    labels, n_child, centres = ops.kmeans(V.contiguous().view(B * N, k),
                                          None if eigvals is None else eigvals.contiguous(),
                                          None if weight is None else weight.contiguous().view(-1),
                                          None if init is None else init.contiguous(), B, N, nk,
                                          float(eigenvalue_threshold), int(max_iter), None)
    return labels.view(B, N).long(), n_child, centres


def pool(x: torch.Tensor, labels: torch.Tensor, K: int):
    """Cluster-mean pooling: x [B, N, D], labels [B, N] int64 -> (pooled [B, K, D] fp32, counts [B, K] int32)."""
    return ops.pool(x.contiguous(), labels.contiguous(), int(K))
