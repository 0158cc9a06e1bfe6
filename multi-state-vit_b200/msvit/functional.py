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
    eigvals: torch.Tensor           # [B, P, k] fp32 (P parents per image)
    n_child: torch.Tensor           # [B, P] int32 children per parent
    iters: torch.Tensor             # [B, P] int32 eigensolver iterations
    affinity: Optional[torch.Tensor] = None  # [B, N, N] fp32 (single-parent case with N % 4 == 0 only)
    verdict: Optional[torch.Tensor] = None   # [B, P] int32, fused path: 1 = the leading block met the tolerance
    iter_cap: int = 0                        # two-kernel path: the iteration cap (stopping there = not converged)
    # eigenvectors / degrees as the kernels wrote them: token order for one parent, segment order (rows sorted by
    # parent) otherwise -- then `perm` maps sorted row -> token row and `token_order` are the plan's [rows, k] / [rows]
    # destination buffers.  The un-permutation runs when `eigvecs` / `degree` is first read, not in every step.
    eigvecs_raw: Optional[torch.Tensor] = None
    degree_raw: Optional[torch.Tensor] = None
    perm: Optional[torch.Tensor] = None
    token_order: Optional[tuple] = None
    _unpermuted: bool = False

    def _to_token_order(self):
        if self.perm is not None and not self._unpermuted:
            idx = self.perm.long()
            self.token_order[0].index_copy_(0, idx, self.eigvecs_raw)
            self.token_order[1].index_copy_(0, idx, self.degree_raw)
            self._unpermuted = True

    @property
    def eigvecs(self) -> torch.Tensor:
        """[B, N, k] fp32, row i = embedding of token i inside its segment."""
        B, N = self.labels.shape
        if self.perm is None:
            return self.eigvecs_raw.view(B, N, -1)
        self._to_token_order()
        return self.token_order[0].view(B, N, -1)

    @property
    def degree(self) -> torch.Tensor:
        """[B, N] fp32 NCut degree of every token inside its segment."""
        B, N = self.labels.shape
        if self.perm is None:
            return self.degree_raw.view(B, N)
        self._to_token_order()
        return self.token_order[1].view(B, N)

    @property
    def converged(self) -> torch.Tensor:
        """[B, P] bool: the wanted eigenpairs met the residual tolerance before the iteration cap.  Derived on demand
        (one small elementwise kernel), so a caller that never asks pays nothing per step."""
        return (self.verdict == 1) if self.verdict is not None else (self.iters < self.iter_cap)


FUSED_BLOCK = 16     # subspace width of the fused kernel
FUSED_MAX_TOKENS = 208


def fused_eligible(N: int, k: int, n_parents: int = 1) -> bool:
    """Shapes msvit_ncut_fused takes: whole images, 16 < N, round16(N) <= 208, ncut_dim <= 12 (the fused kernel always
    iterates on a block of 16 columns, i.e. at least 4 columns of oversampling)."""
    return n_parents == 1 and k + 4 <= FUSED_BLOCK and FUSED_BLOCK < N and ((N + 15) & ~15) <= FUSED_MAX_TOKENS


MAX_DENSE_EIG = 128  # ncut_dim served by the dense solver


def default_block(k: int, oversample: int = 8) -> int:
    """Subspace width of the iterative solvers: k wanted + oversampling, multiple of 4, at most MAX_EIG_BLOCK.
    Returns 0 when k does not fit (ncut_dim > MAX_EIG_BLOCK): such requests go to the dense solver."""
    if k > _lib.MAX_EIG_BLOCK:
        if k > MAX_DENSE_EIG:
            raise ValueError(f"ncut_dim={k} exceeds the supported maximum {MAX_DENSE_EIG}")
        return 0
    m = (k + oversample + 3) & ~3
    m = min(m, _lib.MAX_EIG_BLOCK)
    return max(m, (k + 3) & ~3)


def dense_ncut_eig(A: torch.Tensor, deg: torch.Tensor, k: int):
    """All-eigenpairs solve for large ncut_dim (33 .. 128; the author's logs use 100 on 784 tokens,
    logging/10-03-2024/run_log.png, sandbox/test.py:66): A [S, N, lda] fp32 as written by the affinity kernel,
    deg [S, N] -> (V [S, N, k], lam [S, k]), eigenvalues descending, canonical sign.

    A block subspace iteration converges hopelessly slowly that deep inside the spectrum, so this path does what the
    reference's closed form does (sandbox/test.py:114-118): a full symmetric eigendecomposition of D^-1/2 A D^-1/2.
    It is a LIBRARY call (cuSOLVER through torch.linalg.eigh on the GPU), not one of this repository's kernels; it
    allocates and may synchronise, so plans that use it cannot be captured in a CUDA graph."""
    S, N = deg.shape
    r = torch.rsqrt(deg)
    Abar = A[:, :, :N] * r[:, :, None] * r[:, None, :]
    Abar = 0.5 * (Abar + Abar.transpose(1, 2))
    w, U = torch.linalg.eigh(Abar)
    kk = min(k, N)
    lam = torch.zeros(S, k, dtype=torch.float32, device=A.device)
    V = torch.zeros(S, N, k, dtype=torch.float32, device=A.device)
    lam[:, :kk] = w.flip(-1)[:, :kk]
    Vk = U.flip(-1)[:, :, :kk]
    # largest-|entry| of every column positive, ties -> lowest row
    av = Vk.abs()
    mx = av.max(dim=1, keepdim=True).values
    first = (av >= mx).to(torch.int8).argmax(dim=1)                       # first maximal row
    sgn = torch.sign(torch.gather(Vk, 1, first[:, None, :]))
    sgn = torch.where(sgn == 0, torch.ones_like(sgn), sgn)
    V[:, :, :kk] = Vk * sgn
    return V, lam


class ClusterPlan:
    """Pre-allocated execution plan of the whole hot path for a fixed problem shape.

    All buffers (affinity blocks, degree, eigenvectors, labels, pooled tokens, segment tables) are allocated
    once; `run` only enqueues kernels on the current stream -- no allocation, no host synchronisation -- so a
    plan can be replayed every step and captured in a CUDA graph.  The tensors in the returned ClusterOutput
    are views of plan-owned memory and are overwritten by the next `run`.
    """

    STAGES = ("segments", "affinity", "eig", "kmeans", "compose", "pool")

    def __init__(self, B: int, N: int, D: int, dtype: torch.dtype = torch.float32, device="cuda", *, ncut_dim: int,
                 n_clusters: Optional[int] = None, eigenvalue_threshold: Optional[float] = None, mode: str = "rbf",
                 gamma: float = 3.0, scale: Optional[float] = None, n_parents: int = 1, kmeans_iters: int = 100,
                 eig_iters: int = 60, eig_tol: float = 2e-5, oversample: int = 8, pool_k: Optional[int] = None,
                 want_pool: bool = True, fused: Optional[bool] = None, discretise: str = "kmeans"):
        if mode not in _lib.DIST:
            raise ValueError(f"unknown distance {mode!r}")
        if n_clusters is None and eigenvalue_threshold is None:
            raise ValueError("give n_clusters or eigenvalue_threshold")
        if discretise not in _lib.DISC:
            raise ValueError(f"unknown discretisation {discretise!r} (kmeans | axis_align)")
        self.disc = _lib.DISC[discretise]
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("msvit.cluster_tokens runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dev.index is None:   # "cuda" means the current device; tensors report an explicit index
            dev = torch.device("cuda", torch.cuda.current_device())
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"tokens must be float32 or bfloat16, got {dtype}")
        self.lib = _lib.load()
        self.B, self.N, self.D, self.P = int(B), int(N), int(D), int(n_parents)
        self.dtype, self.device = dtype, dev
        self.dtype_code = _lib.F32 if dtype == torch.float32 else _lib.BF16
        self.k = int(ncut_dim)
        self.block = default_block(self.k, oversample)
        self.dense = self.block == 0   # ncut_dim > 32: dense (library) eigendecomposition
        self.mode = _lib.DIST[mode]
        self.gamma = float(gamma)
        self.scale = float(D) if scale is None else float(scale)
        self.nk = int(n_clusters) if n_clusters is not None else 0
        self.thr = float(eigenvalue_threshold) if eigenvalue_threshold is not None else 0.0
        # eigenpairs that cannot reach the threshold (Ritz value + residual norm < threshold) are never clustered on:
        # they are exempt from the residual test
        self.lam_floor = self.thr if self.nk == 0 else 0.0
        # with a fixed number of clusters K < k the k-means step reads V[:, :K] only: the other pairs need not converge
        self.n_converge = min(self.nk, self.k) if self.nk > 0 else 0
        self.kmeans_iters, self.eig_iters, self.eig_tol = int(kmeans_iters), int(eig_iters), float(eig_tol)
        self.want_pool = bool(want_pool)
        self.Kp = int(pool_k) if pool_k is not None else self.P * (self.nk if self.nk > 0 else min(self.k, _lib.MAX_EIG_BLOCK))
        B, N, D, P, k = self.B, self.N, self.D, self.P, self.k
        self.S = B * P
        # whole images of up to 224 tokens take the fused kernel: the affinity stays in tensor memory
        if self.dense and int(n_parents) != 1:
            raise NotImplementedError("ncut_dim > 32 is served by the dense solver, which takes whole images only")
        can_fuse = fused_eligible(N, self.k, P)
        if fused and not can_fuse:
            raise ValueError("the fused kernel needs whole images (one parent), 16 < tokens <= 208 and ncut_dim <= 12")
        self.fused = can_fuse if fused is None else bool(fused)
        if self.fused:
            self.block = FUSED_BLOCK
        rows = B * N
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            if P == 1:
                self.perm = self.seg_off = self.a_off = self.xs = None
                a_numel = 0 if self.fused else rows * ops.lda_of(N)
            else:
                self.perm = torch.empty(rows, **i32)
                self.seg_off = torch.empty(self.S + 1, **i32)
                self.a_off = torch.empty(self.S + 1, dtype=torch.int64, device=dev)
                self.xs = torch.empty(rows, D, dtype=dtype, device=dev)
                self.V_tok = torch.empty(rows, k, **f32)
                self.deg_tok = torch.empty(rows, **f32)
                a_numel = B * ops.affinity_stride(N)
            self.A = torch.empty(a_numel, **f32)
            self.deg = torch.empty(rows, **f32)
            self.V = torch.empty(rows, k, **f32)
            self.lam = torch.empty(self.S, k, **f32)
            self.iters = torch.empty(self.S, **i32)
            self.labels_sorted = torch.empty(rows, **i32)
            self.n_child = torch.empty(self.S, **i32)
            self.child = torch.empty(B, N, dtype=torch.int64, device=dev)
            self.pooled = torch.empty(B, self.Kp, D, **f32) if want_pool else None
            self.counts = torch.empty(B, self.Kp, **i32) if want_pool else None
            if self.fused:
                self.U = torch.empty(rows, FUSED_BLOCK, **f32)
                self.H = torch.empty(self.S, FUSED_BLOCK * FUSED_BLOCK, **f32)
                self.info = torch.empty(self.S, **i32)

    def run(self, x: torch.Tensor, parent_indices: Optional[torch.Tensor] = None, events=None) -> "ClusterOutput":
        """Enqueue the hot path for x [B, N, D].  `events`, if given, is a list of len(STAGES)+1 CUDA events that
        are recorded around the stages (per-stage timing)."""
        B, N, D, P, k, S = self.B, self.N, self.D, self.P, self.k, self.S
        if tuple(x.shape) != (B, N, D) or x.dtype != self.dtype or x.device != self.device:
            raise ValueError(f"plan was built for {(B, N, D)} {self.dtype} on {self.device}, "
                             f"got {tuple(x.shape)} {x.dtype} on {x.device}")
        if not x.is_contiguous():
            x = x.contiguous()
        if (parent_indices is None) != (P == 1):
            raise ValueError("parent_indices must be given exactly when the plan has n_parents > 1")
        lib, check, p = self.lib, _lib.check, ops._ptr
        rows = B * N
        st = torch.cuda.current_stream(self.device).cuda_stream

        def mark(i):
            if events is not None:
                events[i].record()

        with torch.cuda.device(self.device):
            mark(0)
            if P > 1:
                if parent_indices.shape != (B, N) or parent_indices.dtype != torch.int64:
                    raise ValueError("parent_indices must be int64 [batch, tokens]")
                if parent_indices.device != self.device:
                    raise ValueError(f"parent_indices must live on {self.device} (got {parent_indices.device}); the kernels "
                                     "take raw device pointers")
                parent_indices = parent_indices.contiguous()
                check(lib.msvit_build_segments(p(parent_indices), p(self.perm), p(self.seg_off), p(self.a_off), B, N, P,
                                               st), "msvit_build_segments")
                check(lib.msvit_gather_rows(p(x), self.dtype_code, p(self.perm), p(self.xs), rows, D, st),
                      "msvit_gather_rows")
                xs = self.xs
            else:
                xs = x
            mark(1)
            if self.fused:
                check(lib.msvit_ncut_fused(p(x), self.dtype_code, p(self.deg), p(self.U), p(self.H), p(self.iters),
                                           p(self.info), rows, S, N, D, self.mode, self.gamma, self.scale, self.block,
                                           self.eig_iters, self.eig_tol, self.lam_floor, self.n_converge, st),
                      "msvit_ncut_fused")
                mark(2)
                mark(3)
                check(lib.msvit_ritz_kmeans(p(self.U), p(self.H), p(self.info), p(self.deg), p(self.V), p(self.lam), None,
                                            p(self.child), p(self.n_child), rows, S, N, k, self.block, self.n_converge,
                                            self.nk, self.thr, self.kmeans_iters, self.disc, st), "msvit_ritz_kmeans")
                mark(4)
            else:
                check(lib.msvit_affinity_degree(p(xs), self.dtype_code, p(self.A), p(self.deg), rows, S, N, D, self.mode,
                                                self.gamma, self.scale, p(self.seg_off), p(self.a_off), st),
                      "msvit_affinity_degree")
                mark(2)
                if self.dense:
                    Vd, lamd = dense_ncut_eig(self.A.view(B, N, ops.lda_of(N)), self.deg.view(B, N), k)
                    self.V.copy_(Vd.view(rows, k))
                    self.lam.copy_(lamd)
                    self.iters.fill_(1)
                else:
                    check(lib.msvit_ncut_eig(p(self.A), p(self.deg), p(self.V), p(self.lam), p(self.iters), rows, S, N, k,
                                             self.block, self.eig_iters, self.eig_tol, self.lam_floor, self.n_converge,
                                             p(self.seg_off), p(self.a_off), st), "msvit_ncut_eig")
                mark(3)
                check(lib.msvit_discretise(p(self.V), p(self.lam), p(self.deg), None, p(self.labels_sorted),
                                           p(self.n_child), None, rows, S, N, k, self.nk, self.thr, self.kmeans_iters,
                                           self.disc, p(self.seg_off), st), "msvit_discretise")
                mark(4)
                check(lib.msvit_compose_labels(p(self.labels_sorted), p(self.n_child), p(self.perm), p(self.seg_off),
                                               p(self.child), B, N, P, st), "msvit_compose_labels")
            mark(5)
            if self.want_pool:
                check(lib.msvit_pool(p(x), self.dtype_code, p(self.child), p(self.pooled), p(self.counts), B, N, D,
                                     self.Kp, st), "msvit_pool")
            mark(6)

        aff = self.A.view(B, N, N) if (P == 1 and N % 4 == 0 and not self.fused) else None
        # the fused kernel reports its verdict; the two-kernel solver stops at the cap only when it did not converge
        return ClusterOutput(labels=self.child, pooled=self.pooled, counts=self.counts,
                             eigvals=self.lam.view(B, P, k), n_child=self.n_child.view(B, P),
                             iters=self.iters.view(B, P), affinity=aff,
                             verdict=self.info.view(B, P) if self.fused else None, iter_cap=self.eig_iters,
                             eigvecs_raw=self.V, degree_raw=self.deg, perm=self.perm,
                             token_order=(self.V_tok, self.deg_tok) if self.perm is not None else None)


def cluster_tokens(x: torch.Tensor, parent_indices: Optional[torch.Tensor] = None, *, ncut_dim: int,
                   n_clusters: Optional[int] = None, eigenvalue_threshold: Optional[float] = None,
                   mode: str = "rbf", gamma: float = 3.0, scale: Optional[float] = None,
                   n_parents: Optional[int] = None, kmeans_iters: int = 100, eig_iters: int = 60,
                   eig_tol: float = 2e-5, oversample: int = 8, pool_k: Optional[int] = None,
                   want_pool: bool = True, keep_affinity: bool = False, fused: Optional[bool] = None,
                   discretise: str = "kmeans") -> ClusterOutput:
    """One-shot form: builds a ClusterPlan for x's shape and runs it (buffers are owned by the result)."""
    if x.dim() != 3:
        raise ValueError("x must be [batch, tokens, hidden]")
    if not x.is_cuda:
        raise RuntimeError("msvit.cluster_tokens runs on CUDA (sm_100a) only; there is no CPU fallback")
    B, N, D = x.shape
    if parent_indices is None:
        P = 1
    else:
        if tuple(parent_indices.shape) != (B, N):
            raise ValueError("parent_indices must be [batch, tokens]")
        if parent_indices.dtype != torch.int64:
            raise ValueError(f"parent_indices must be int64 (got {parent_indices.dtype})")
        if parent_indices.device != x.device:
            parent_indices = parent_indices.to(x.device)   # e.g. the caller's initial all-zero indices built on the host
        # the reference reads this on the host too (modeling_spectral.py:80); pass n_parents to skip the sync
        P = int(n_parents) if n_parents is not None else int(parent_indices.max().item()) + 1
        if P == 1:
            parent_indices = None
    plan = ClusterPlan(B, N, D, x.dtype, x.device, ncut_dim=ncut_dim, n_clusters=n_clusters,
                       eigenvalue_threshold=eigenvalue_threshold, mode=mode, gamma=gamma, scale=scale, n_parents=P,
                       kmeans_iters=kmeans_iters, eig_iters=eig_iters, eig_tol=eig_tol, oversample=oversample,
                       pool_k=pool_k, want_pool=want_pool, fused=False if keep_affinity else fused,
                       discretise=discretise)
    out = plan.run(x, parent_indices)
    if not keep_affinity:
        out.affinity = None
    return out


def affinity(x: torch.Tensor, mode: str = "rbf", gamma: float = 3.0, scale: Optional[float] = None):
    """x [B, N, D] -> (A [B, N, lda] fp32 (lda = N rounded up to 4, pad columns are 0), deg [B, N])."""
    B, N, D = x.shape
    x = x.contiguous()
    s = float(D) if scale is None else float(scale)
    lda = ops.lda_of(N)
    A, deg = ops.affinity_degree(x.view(B * N, D), B, N, _lib.DIST[mode], float(gamma), s, None, None, B * N * lda, True)
    return A.view(B, N, lda), deg.view(B, N)


def ncut_eig(A: torch.Tensor, deg: torch.Tensor, k: int, *, max_iter: int = 60, tol: float = 2e-5, oversample: int = 8,
             lam_floor: float = 0.0, n_converge: int = 0):
    """A [B, N, lda] (as returned by `affinity`), deg [B, N] -> (V [B, N, k], lam [B, k], iters [B])."""
    B, N, lda = A.shape
    if lda != ops.lda_of(N):
        raise ValueError("A must be [B, N, (N+3)&~3]")
    if default_block(int(k), oversample) == 0:
        V, lam = dense_ncut_eig(A.contiguous(), deg.contiguous(), int(k))
        return V, lam, torch.ones(B, dtype=torch.int32, device=A.device)
    V, lam, iters = ops.ncut_eig(A.contiguous().view(-1), deg.contiguous().view(-1), B, N, int(k),
                                 default_block(int(k), oversample), int(max_iter), float(tol), float(lam_floor),
                                 int(n_converge), None, None)
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


def attention_mask(cluster_indices: torch.Tensor, max_n_clusters: Optional[int] = None) -> torch.Tensor:
    """Cluster-restricted attention mask of the multi-state encoder (modeling_msvitencoder.py:426-467):
    cluster_indices [B, N] int64 -> bool [B, 1, L, L], L = 2C + N, sequence [T_0, R_0, .., T_{C-1}, R_{C-1}, tokens].

    C = max_n_clusters; when it is not given it is read from the labels on the host, as the reference does
    (`torch.max(cluster_indices).item() + 1`, :451) -- pass it (e.g. the plan's cluster bound) to stay asynchronous."""
    if cluster_indices.dim() != 2 or cluster_indices.dtype != torch.int64:
        raise ValueError("cluster_indices must be int64 [batch, tokens]")
    if not cluster_indices.is_cuda:
        raise RuntimeError("msvit.attention_mask runs on CUDA (sm_100a) only; there is no CPU fallback")
    cluster_indices = cluster_indices.contiguous()
    B, N = cluster_indices.shape
    C = int(max_n_clusters) if max_n_clusters is not None else int(cluster_indices.max().item()) + 1
    L = 2 * C + N
    with torch.cuda.device(cluster_indices.device):
        mask = torch.empty(B, 1, L, L, dtype=torch.uint8, device=cluster_indices.device)
        st = torch.cuda.current_stream(cluster_indices.device).cuda_stream
        _lib.check(_lib.load().msvit_attention_mask(ops._ptr(cluster_indices), ops._ptr(mask), B, N, C, st),
                   "msvit_attention_mask")
    return mask.view(torch.bool)


def cluster_attention_stats(attention_probs: torch.Tensor, cluster_indices: torch.Tensor, n_clusters: int):
    """Cluster-compressed attention statistics of compress_tokens_with_cluster_indices (msvitencoder.py:182-190):

        transmitter [B, H, N, C] = sum of attention_probs[b, h, q, :] over the KEYS of each cluster
        receiver    [B, H, C, N] = mean of attention_probs[b, h, :, k] over the QUERIES of each cluster (empty -> 0)

    attention_probs [B, H, N, N] fp32, cluster_indices [B, N] int64, n_clusters <= 64 (the author's logs show up to
    39 clusters per image); the reference materialises a 5-D broadcast.  An empty cluster gives receiver rows of 0
    where the reference's 0/0 gives NaN."""
    if attention_probs.dim() != 4 or attention_probs.shape[-1] != attention_probs.shape[-2]:
        raise ValueError("attention_probs must be [batch, heads, tokens, tokens]")
    if not attention_probs.is_cuda:
        raise RuntimeError("msvit.cluster_attention_stats runs on CUDA (sm_100a) only; there is no CPU fallback")
    if attention_probs.dtype != torch.float32:
        raise TypeError("attention_probs must be float32")
    B, H, N, _ = attention_probs.shape
    if tuple(cluster_indices.shape) != (B, N) or cluster_indices.dtype != torch.int64:
        raise ValueError("cluster_indices must be int64 [batch, tokens]")
    if cluster_indices.device != attention_probs.device:
        raise ValueError("cluster_indices must live on the same device as attention_probs")
    C = int(n_clusters)
    a = attention_probs.contiguous()
    lab = cluster_indices.contiguous()
    with torch.cuda.device(a.device):
        tr = torch.empty(B, H, N, C, dtype=torch.float32, device=a.device)
        st = torch.cuda.current_stream(a.device).cuda_stream
        if N <= 256 and C <= 16:
            # one pass over the attention tensor for both statistics (the faster route in this range)
            rc = torch.empty(B, H, C, N, dtype=torch.float32, device=a.device)
            _lib.check(_lib.load().msvit_cluster_attention_stats(ops._ptr(a), ops._ptr(lab), ops._ptr(tr), ops._ptr(rc),
                                                                 B, H, N, C, st), "msvit_cluster_attention_stats")
            return tr, rc
        # longer rows: key sums, then the query means as cluster-mean pooling of the [B*H, N, N] view (two passes)
        _lib.check(_lib.load().msvit_cluster_key_sums(ops._ptr(a), ops._ptr(lab), ops._ptr(tr), B, H, N, C, st),
                   "msvit_cluster_key_sums")
        lab_h = lab[:, None, :].expand(B, H, N).reshape(B * H, N).contiguous()
        rc, _ = ops.pool(a.view(B * H, N, N), lab_h, C)
    return tr, rc.view(B, H, C, N)


@dataclass
class HostResult:
    labels: torch.Tensor   # [B, N] int64, pinned host memory
    pooled: torch.Tensor   # [B, K, D] fp32, pinned host memory
    counts: torch.Tensor   # [B, K] int32, pinned host memory


class HostClusterer:
    """End-to-end form of the hot path for tokens that live in HOST memory.

    `run(x_host)` streams the batch to the GPU in chunks over `n_streams` CUDA streams -- pinned host tokens ->
    H2D copy -> the kernel sequence of ClusterPlan -> D2H copy of labels, pooled tokens and counts -- so that the
    copies of one chunk overlap the kernels of another.  Everything (device buffers, plans, pinned result buffers)
    is allocated once; `run` returns after all chunks have landed in the pinned result buffers.
    """

    def __init__(self, B: int, N: int, D: int, dtype: torch.dtype = torch.float32, device="cuda", *, chunk: int = 128,
                 n_streams: int = 3, **plan_kwargs):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("msvit.HostClusterer runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.B, self.N, self.D, self.dtype, self.device = int(B), int(N), int(D), dtype, dev
        self.chunk = max(1, min(int(chunk), self.B))
        self.bounds = [(b0, min(self.B, b0 + self.chunk)) for b0 in range(0, self.B, self.chunk)]
        n_streams = max(1, min(int(n_streams), len(self.bounds)))
        if plan_kwargs.get("n_parents", 1) != 1:
            raise ValueError("HostClusterer clusters whole images (single parent)")
        self.streams, self.plans, self.xdev = [], [], []
        self.tail_plan = self.tail_x = None
        with torch.cuda.device(dev):
            for _ in range(n_streams):
                self.streams.append(torch.cuda.Stream(dev))
                self.plans.append(ClusterPlan(self.chunk, N, D, dtype, dev, **plan_kwargs))
                self.xdev.append(torch.empty(self.chunk, N, D, dtype=dtype, device=dev))
            tail = self.bounds[-1][1] - self.bounds[-1][0]
            if tail != self.chunk:
                self.tail_plan = ClusterPlan(tail, N, D, dtype, dev, **plan_kwargs)
                self.tail_x = torch.empty(tail, N, D, dtype=dtype, device=dev)
            Kp = self.plans[0].Kp
            self.labels = torch.empty(B, N, dtype=torch.int64).pin_memory()
            self.pooled = torch.empty(B, Kp, D, dtype=torch.float32).pin_memory()
            self.counts = torch.empty(B, Kp, dtype=torch.int32).pin_memory()
        self.h2d_bytes = B * N * D * (4 if dtype == torch.float32 else 2)
        self.d2h_bytes = self.labels.numel() * 8 + self.pooled.numel() * 4 + self.counts.numel() * 4

    def run(self, x_host: torch.Tensor) -> HostResult:
        if x_host.is_cuda:
            raise ValueError("HostClusterer.run takes host tokens; use cluster_tokens / ClusterPlan for device tensors")
        if tuple(x_host.shape) != (self.B, self.N, self.D) or x_host.dtype != self.dtype:
            raise ValueError(f"expected host tokens {(self.B, self.N, self.D)} {self.dtype}")
        if not x_host.is_pinned():
            x_host = x_host.pin_memory()
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        for i, (b0, b1) in enumerate(self.bounds):
            j = i % len(self.streams)
            st = self.streams[j]
            if i < len(self.streams):
                st.wait_event(start)
            full = (b1 - b0) == self.chunk
            plan = self.plans[j] if full else self.tail_plan
            xd = self.xdev[j] if full else self.tail_x
            with torch.cuda.stream(st):
                xd.copy_(x_host[b0:b1], non_blocking=True)
                out = plan.run(xd)
                self.labels[b0:b1].copy_(out.labels, non_blocking=True)
                self.pooled[b0:b1].copy_(out.pooled, non_blocking=True)
                self.counts[b0:b1].copy_(out.counts, non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        cur.synchronize()
        return HostResult(self.labels, self.pooled, self.counts)
