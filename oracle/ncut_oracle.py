"""CPU oracle for the token-grouping hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

    PARITY: pinned where the reference is executable, unpinned where it is not.  Four pieces of the reference's own
    plain-torch arithmetic are executed from the checkout by tests/golden/make_reference_fixtures.py and this file
    reproduces them (tests/test_reference_fixtures.py): the attention mask (modeling_msvitencoder.py:426-467), the
    transmitter / receiver attention statistics (:169,182-190), the closed-form NCut of sandbox/test.py:100,106-118
    (normprod distance, exp, degree, normalised Laplacian, eigh -- to 1e-9) and the per-label mean centres /
    nearest-centre assignment (modeling_spectral.py:125-127,129).  What stays UNPINNED is the third-party arithmetic
    absent from /root/reference and from this image (ncut-pytorch==1.7.9, cuml~=24.10, fast_pytorch_kmeans;
    requirements.txt:27,32): the solver / sampling / sign conventions of NCUT.fit_transform, cuML's k-means
    initialisation and kway_ncut are restated from the call sites and the published algorithms.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (multi-state-vit_b200/) never does.

What each function follows (paths relative to /root/reference):

  pairwise_distance / affinity  sandbox/test.py:108-112 (normprod distance then exp(-d/gamma));
                                sandbox/ncut_euclidean.py:19,23-29 (rbf on raw M == cosine on
                                normalised M  =>  d_rbf = 1/2 |xi-xj|^2);
                                model/clustering/modeling_spectral.py:54-61 (gamma = 3.0, rbf|cosine)
  ncut_eig                      sandbox/test.py:114-118 (deg, I - D^-1/2 A D^-1/2, eigh, leading k)
  select_n_children             model/clustering/modeling_spectral.py:87,92-93
  kway_ncut                     model/clustering/modeling_spectral.py:136-138, modeling_axisalign.py:35-36 (third-party
                                algorithm, restated from Yu & Shi 2003; unpinned)
  kmeans                        model/clustering/modeling_spectral.py:90 (Lloyd on V[:, :K]),
                                :129 (assignment = argmin cdist), :125-133,271-278 (centre = label mean,
                                centroid-seeded k-means)
  pool                          model/clustering/modeling_spectral.py:125-127,271-273
  cluster_tokens                model/clustering/modeling_spectral.py:72-94 (per-parent loop, offset labels),
                                :260-279 (per-image variant), multistate_encoder/modeling_msvitencoder.py:491-499
                                (label ordering the caller relies on)
  global_kmeans                 model/clustering/modeling_spectral.py:254-256 (flattened-batch clustering)

Everything is deterministic: exact `eigh`, farthest-point k-means initialisation,
lowest-index tie breaks, first-occurrence relabelling.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

DIST_MODES = ("rbf", "cosine", "normprod")


# --------------------------------------------------------------------------- affinity
def pairwise_distance(x: torch.Tensor, mode: str = "rbf", scale: Optional[float] = None) -> torch.Tensor:
    """x [n, D] -> d [n, n] >= 0.

    rbf      : d = 1/2 |xi - xj|^2 / s          (sandbox/ncut_euclidean.py:19,23-29)
    cosine   : d = 1 - cos(xi, xj)              (modeling_spectral.py:62-69)
    normprod : d = (|xi||xj| - xi.xj) / s       (sandbox/test.py:108-110)
    s defaults to D for rbf/normprod (resolution independent gamma) and is unused for cosine.
    """
    if mode not in DIST_MODES:
        raise ValueError(f"unknown distance mode {mode!r}")
    n, D = x.shape
    g = x @ x.T
    sq = (x * x).sum(-1)
    if mode == "rbf":
        s = float(D) if scale is None else float(scale)
        d = (0.5 * (sq[:, None] + sq[None, :]) - g) / s
    elif mode == "cosine":
        rn = torch.rsqrt(torch.clamp_min(sq, 1e-30))
        d = 1.0 - g * rn[:, None] * rn[None, :]
    else:
        s = float(D) if scale is None else float(scale)
        nrm = torch.sqrt(sq)
        d = (nrm[:, None] * nrm[None, :] - g) / s
    return torch.clamp_min(d, 0.0)


def affinity(x: torch.Tensor, mode: str = "rbf", gamma: float = 3.0, scale: Optional[float] = None) -> torch.Tensor:
    """A = exp(-d / gamma)   (sandbox/test.py:112; gamma = affinity_focal_gamma, modeling_spectral.py:59)."""
    return torch.exp(-pairwise_distance(x, mode, scale) / gamma)


# --------------------------------------------------------------------------- eigensolve
def sign_fix(V: torch.Tensor) -> torch.Tensor:
    """Make the largest-|entry| of every column positive (ties -> lowest row)."""
    if V.numel() == 0:
        return V
    idx = torch.argmax(V.abs(), dim=0)  # first maximal index
    sgn = torch.sign(V[idx, torch.arange(V.shape[1])])
    sgn = torch.where(sgn == 0, torch.ones_like(sgn), sgn)
    return V * sgn[None, :]


def ncut_eig(A: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """A [n, n] symmetric affinity -> (V [n, k], lam [k], deg [n]).

    deg = A 1 ; Abar = D^-1/2 A D^-1/2 ; top-k eigenpairs of Abar (== smallest k of
    L = I - Abar, sandbox/test.py:114-118), eigenvalues descending, sign-fixed columns.
    If k > n the trailing columns / eigenvalues are zero.
    """
    n = A.shape[0]
    deg = A.sum(-1)
    r = torch.rsqrt(deg)
    Abar = A * r[:, None] * r[None, :]
    Abar = 0.5 * (Abar + Abar.T)
    w, U = torch.linalg.eigh(Abar)
    kk = min(k, n)
    lam = torch.flip(w, dims=[0])[:kk]
    V = torch.flip(U, dims=[1])[:, :kk]
    V = sign_fix(V)
    if kk < k:
        lam = torch.cat([lam, lam.new_zeros(k - kk)])
        V = torch.cat([V, V.new_zeros(n, k - kk)], dim=1)
    return V, lam, deg


def select_n_children(lam: torch.Tensor, threshold: float) -> int:
    """n_child = #{lam > thr}; no eigenvalue above threshold -> one child (modeling_spectral.py:87,92-93)."""
    return max(1, int((lam > threshold).sum().item()))


# --------------------------------------------------------------------------- k-means
def _sqdist(P: torch.Tensor, C: torch.Tensor) -> torch.Tensor:
    return ((P[:, None, :] - C[None, :, :]) ** 2).sum(-1)


def farthest_point_init(P: torch.Tensor, K: int, weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Deterministic seeding: first centre = row argmax(weight) (row 0 if no weight), then
    repeatedly the row farthest from the chosen set (ties -> lowest row)."""
    n = P.shape[0]
    first = int(torch.argmax(weight).item()) if weight is not None else 0
    chosen = [first]
    mind = ((P - P[first]) ** 2).sum(-1)
    for _ in range(1, K):
        nxt = int(torch.argmax(mind).item())
        chosen.append(nxt)
        mind = torch.minimum(mind, ((P - P[nxt]) ** 2).sum(-1))
    return P[torch.tensor(chosen, dtype=torch.long)].clone()


def canonical_relabel(labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rename clusters in order of first occurrence -> (labels', old_id_of_new)."""
    lab = labels.tolist()
    mapping = {}
    out = []
    for v in lab:
        if v not in mapping:
            mapping[v] = len(mapping)
        out.append(mapping[v])
    order = sorted(mapping, key=mapping.get)
    return torch.tensor(out, dtype=torch.long), torch.tensor(order, dtype=torch.long)


def kmeans(P: torch.Tensor, K: int, init: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None,
           iters: int = 100) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Lloyd k-means on rows of P [n, d] -> (labels [n] int64 canonical, centres [C, d], C).

    assignment = argmin_c |p - mu_c|^2, ties -> lowest c (modeling_spectral.py:129);
    update = label mean, empty cluster keeps its previous centre (:125-127);
    stop when labels stop changing or after `iters` assignments;
    C = number of non-empty clusters after canonical relabelling.
    """
    n = P.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.long), P.new_zeros(0, P.shape[1]), 0
    K = max(1, min(int(K), n))
    C = init.clone().to(P.dtype) if init is not None else farthest_point_init(P, K, weight)
    K = C.shape[0]
    labels = torch.full((n,), -1, dtype=torch.long)
    for _ in range(max(1, iters)):
        new = torch.argmin(_sqdist(P, C), dim=1)
        if torch.equal(new, labels):
            break
        labels = new
        for c in range(K):
            m = labels == c
            if m.any():
                C[c] = P[m].mean(0)
    labels, order = canonical_relabel(labels)
    return labels, C[order], int(order.numel())


# --------------------------------------------------------------------------- axis-aligned discretisation
def kway_ncut(V: torch.Tensor, K: int, weight: Optional[torch.Tensor] = None, iters: int = 100):
    """Axis-aligned discretisation of a spectral embedding: rows of V[:, :K] -> (labels [n] canonical, C, R [K, K]).

    UNPINNED (third-party algorithm): the reference calls ncut_pytorch.kway_ncut
    (model/clustering/modeling_spectral.py:136-138, model/clustering/modeling_axisalign.py:35-36), which is absent from
    the checkout; this restates the published algorithm it implements, Yu & Shi, "Multiclass spectral clustering"
    (ICCV 2003): scale the rows to unit length, pick K rows greedily as orthogonal as possible for the initial
    rotation, then alternate  labels = argmax(Xn R)  and  R = V U^T  with  onehot(labels)^T Xn = U S V^T.
    Deterministic choices of this repository: first row = argmax weight (row 0 without weights), ties -> lowest index,
    stop when the labels stop changing, first-occurrence relabelling.
    """
    n = V.shape[0]
    K = max(1, min(int(K), n, V.shape[1]))
    X = V[:, :K]
    nrm = X.norm(dim=1, keepdim=True)
    Xn = torch.where(nrm > 0, X / nrm.clamp_min(1e-300), torch.zeros_like(X))
    first = int(torch.argmax(weight).item()) if weight is not None else 0
    R = torch.zeros(K, K, dtype=X.dtype)
    R[:, 0] = Xn[first]
    c = torch.zeros(n, dtype=X.dtype)
    for j in range(1, K):
        c = c + (Xn @ R[:, j - 1]).abs()
        R[:, j] = Xn[int(torch.argmin(c).item())]
    labels = torch.full((n,), -1, dtype=torch.long)
    for _ in range(max(1, iters)):
        new = torch.argmax(Xn @ R, dim=1)
        if torch.equal(new, labels):
            break
        labels = new
        M = torch.zeros(K, K, dtype=X.dtype).index_add_(0, labels, Xn)      # onehot(labels)^T Xn
        U, S, Vh = torch.linalg.svd(M)
        R = Vh.T @ U.T
    labels, order = canonical_relabel(labels)
    return labels, int(order.numel()), R


# --------------------------------------------------------------------------- pooling
def pool(x: torch.Tensor, labels: torch.Tensor, K: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [B, N, D], labels [B, N] -> pooled [B, K, D] (cluster means, empty -> 0), counts [B, K] int32.
    Labels outside [0, K) are ignored."""
    B, N, D = x.shape
    pooled = x.new_zeros(B, K, D)
    counts = torch.zeros(B, K, dtype=torch.int32)
    for b in range(B):
        for c in range(K):
            m = labels[b] == c
            cnt = int(m.sum().item())
            counts[b, c] = cnt
            if cnt:
                pooled[b, c] = x[b][m].mean(0)
    return pooled, counts


# --------------------------------------------------------------------------- whole path
def cluster_segment(xs: torch.Tensor, k: int, n_clusters: Optional[int], threshold: Optional[float],
                    mode: str, gamma: float, scale: Optional[float], kmeans_iters: int, discretise: str = "kmeans"):
    """One (image, parent) segment xs [n, D] -> (labels_local [n], C, V [n,k], lam [k])."""
    n = xs.shape[0]
    A = affinity(xs, mode, gamma, scale)
    V, lam, deg = ncut_eig(A, k)
    if n_clusters is not None:
        K = int(n_clusters)
    else:
        K = select_n_children(lam, float(threshold))
    K = max(1, min(K, n, k))
    if discretise == "axis_align":
        labels, C, _ = kway_ncut(V, K, weight=deg, iters=kmeans_iters)
    else:
        labels, _, C = kmeans(V[:, :K], K, weight=deg, iters=kmeans_iters)
    return labels, C, V, lam


def cluster_tokens(x: torch.Tensor, parent_indices: Optional[torch.Tensor] = None, *, ncut_dim: int,
                   n_clusters: Optional[int] = None, eigenvalue_threshold: Optional[float] = None,
                   mode: str = "rbf", gamma: float = 3.0, scale: Optional[float] = None,
                   kmeans_iters: int = 100, discretise: str = "kmeans"):
    """Per-(image, parent) NCut clustering.

    x [B, N, D]; parent_indices [B, N] int64 (None = all zeros).
    Returns child_indices [B, N] int64 with the caller's contract (msvitencoder.py:491-499):
    per image contiguous ids, children of parent p in a contiguous range, ranges ordered by p;
    plus eigvecs [B, N, k] (row i = embedding of token i inside its own segment) and
    eigvals [B, P, k] (P = max parents per image).
    """
    B, N, D = x.shape
    if parent_indices is None:
        parent_indices = torch.zeros(B, N, dtype=torch.long)
    P = int(parent_indices.max().item()) + 1
    child = torch.zeros(B, N, dtype=torch.long)
    eigvecs = x.new_zeros(B, N, ncut_dim)
    eigvals = x.new_zeros(B, P, ncut_dim)
    n_children = torch.zeros(B, P, dtype=torch.long)
    for b in range(B):
        offset = 0
        for p in range(P):
            idx = torch.nonzero(parent_indices[b] == p).flatten()
            if idx.numel() == 0:
                continue
            labels, C, V, lam = cluster_segment(x[b, idx], ncut_dim, n_clusters, eigenvalue_threshold,
                                                mode, gamma, scale, kmeans_iters, discretise)
            child[b, idx] = offset + labels
            eigvecs[b, idx] = V
            eigvals[b, p] = lam
            n_children[b, p] = C
            offset += C
    return child, eigvecs, eigvals, n_children


# --------------------------------------------------------------------------- flattened-batch Nystrom NCut
def nystrom_ncut(x: torch.Tensor, k: int, sample_idx: torch.Tensor, knn: int = 10, mode: str = "rbf", gamma: float = 3.0,
                 scale: Optional[float] = None):
    """Restates the sample -> exact NCut -> kNN propagation scheme of ncut_pytorch.NCUT.fit_transform for inputs larger
    than `num_sample` (call sites model/clustering/modeling_spectral.py:254-256, modeling_fps.py:36-37).  UNPINNED: the
    package is absent; this mirrors multi-state-vit_b200/msvit/nystrom.py step by step in fp64 with exact eigh.
    x [n, D], sample_idx sorted row ids -> (eigvecs [n, k], eigvals [k])."""
    n, D = x.shape
    xs = x[sample_idx]
    A = affinity(xs, mode, gamma, scale)
    Vs, lam, _ = ncut_eig(A, k)
    if sample_idx.numel() == n:
        return Vs, lam
    out = x.new_zeros(n, k)
    out[sample_idx] = Vs
    rest = torch.ones(n, dtype=torch.bool)
    rest[sample_idx] = False
    ridx = torch.nonzero(rest).flatten()
    s = float(D) if scale is None else float(scale)
    xr = x[ridx]
    g = xr @ xs.T
    if mode == "cosine":
        d = 1.0 - g * torch.rsqrt((xr * xr).sum(-1))[:, None] * torch.rsqrt((xs * xs).sum(-1))[None, :]
    elif mode == "rbf":
        d = (0.5 * ((xr * xr).sum(-1)[:, None] + (xs * xs).sum(-1)[None, :]) - g) / s
    else:
        d = (xr.norm(dim=-1)[:, None] * xs.norm(dim=-1)[None, :] - g) / s
    a = torch.exp(-torch.clamp_min(d, 0.0) / gamma)
    w, nb = torch.topk(a, min(knn, xs.shape[0]), dim=1)
    w = w / w.sum(dim=1, keepdim=True)
    out[ridx] = (w[:, :, None] * Vs[nb]).sum(dim=1)
    return out, lam


# --------------------------------------------------------------------------- dataset-level k-means
def global_kmeans(feats: torch.Tensor, k: int, iters: int, init: Optional[torch.Tensor] = None):
    """DeepCluster-style Lloyd over all rows (modeling_spectral.py:254-256 flattened batch).
    init = first k rows unless given; fixed `iters` iterations of assign -> update; empty keeps centre.
    Returns (centres [k, D], labels [n] of the LAST assignment, counts [k])."""
    C = (feats[:k] if init is None else init).clone()
    n = feats.shape[0]
    labels = torch.zeros(n, dtype=torch.long)
    counts = torch.zeros(k, dtype=torch.long)
    for _ in range(iters):
        cn = (C * C).sum(-1)
        score = feats @ C.T - 0.5 * cn[None, :]  # argmax score == argmin |f - c|^2
        labels = torch.argmax(score, dim=1)
        sums = torch.zeros_like(C).index_add_(0, labels, feats)
        counts = torch.bincount(labels, minlength=k)
        nz = counts > 0
        C[nz] = sums[nz] / counts[nz].to(C.dtype)[:, None]
    return C, labels, counts


# --------------------------------------------------------------------------- attention mask of the ms-ViT caller
def attention_mask(cluster_indices: torch.Tensor) -> torch.Tensor:
    """Restates MultiStateViTEncoderBackbone._construct_attention_mask
    (model/multistate_encoder/modeling_msvitencoder.py:426-467): cluster_indices [B, N] -> bool [B, 1, L, L] with
    L = 2C + N, C = global max cluster count, sequence [T_0, R_0, .., T_{C-1}, R_{C-1}, tokens]."""
    B, N = cluster_indices.shape
    n_clusters = cluster_indices.max(dim=1).values + 1
    C = int(n_clusters.max())
    L = 2 * C + N
    mask = torch.zeros(B, L, L, dtype=torch.bool)
    tok = slice(2 * C, L)
    mask[:, tok, tok] = cluster_indices[:, :, None] == cluster_indices[:, None, :]          # same cluster (:433-436)
    member = torch.arange(C)[None, :, None] == cluster_indices[:, None, :]                    # [B, C, N] (:438-440)
    mask[:, 0:2 * C:2, tok] = member                                                         # T_c -> its tokens (:442)
    mask[:, tok, 1:2 * C:2] = member.transpose(1, 2)                                         # token -> R_c (:444)
    live = torch.arange(C)[None, :] < n_clusters[:, None]                                    # [B, C] (:447-450)
    mask[:, 1:2 * C:2, 0:2 * C:2] = live[:, :, None] & live[:, None, :]                      # R_r -> T_t (:451)
    return mask[:, None]


def cluster_attention_stats(attention_probs: torch.Tensor, cluster_indices: torch.Tensor, n_clusters: int):
    """Restates the transmitter / receiver statistics of compress_tokens_with_cluster_indices
    (model/multistate_encoder/modeling_msvitencoder.py:182-190): one-hot masks [B, N, C]; transmitter = attention
    summed over the keys of each cluster [B, H, N, C]; receiver = attention averaged over the queries of each
    cluster [B, H, C, N] (an empty cluster gives 0 here, 0/0 in the reference)."""
    masks = (cluster_indices[..., None] == torch.arange(n_clusters)).to(attention_probs.dtype)     # [B, N, C]
    transmitter = torch.einsum("bhqk,bkc->bhqc", attention_probs, masks)
    counts = masks.sum(1)                                                                          # [B, C]
    receiver = torch.einsum("bhqk,bqc->bhck", attention_probs, masks) / counts.clamp(min=1)[:, None, :, None]
    return transmitter, receiver


# --------------------------------------------------------------------------- helpers for tests
def round_to_bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(x.dtype)


def truncate_to_tf32(x: torch.Tensor) -> torch.Tensor:
    """Drop the low 13 mantissa bits (what the tensor core reads from an fp32 operand)."""
    xi = x.to(torch.float32).contiguous().view(torch.int32)
    return (xi & ~0x1FFF).view(torch.float32).to(x.dtype)


def round_to_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round to nearest TF32 (10 explicit mantissa bits), ties away from zero: what the affinity kernel feeds the
    tensor core for an fp32 operand."""
    xi = x.to(torch.float32).contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)


def subspace_distance(V: torch.Tensor, W: torch.Tensor) -> float:
    """|V V^T - W W^T|_F for orthonormal column blocks."""
    return float(torch.linalg.norm(V @ V.T - W @ W.T).item())
