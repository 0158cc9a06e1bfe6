"""Flattened-batch ("dataset-level") NCut with Nystrom-style propagation -- SURVEY.md section 8(f).4.

Reference call sites: model/clustering/modeling_spectral.py:254-256 (`self.ncut.fit_transform(x.flatten(0, -2))` over all
B*N tokens, `num_sample=10000`, :57) and model/clustering/modeling_fps.py:36-37.  The arithmetic lives in
ncut-pytorch==1.7.9 (requirements.txt:27), which is absent from the reference checkout and from this image, so what
follows restates the published scheme and is UNPINNED:

    1. sample at most `num_sample` rows (seeded random permutation, `sample_method="random"` at modeling_spectral.py:56);
    2. exact NCut on the sample: affinity + degree (this repository's tcgen05 kernel), leading eigenvectors of
       D^-1/2 A D^-1/2 (this repository's subspace-iteration kernel up to 1024 sampled rows and ncut_dim <= 24; above
       that a torch block iteration on cuBLAS GEMMs -- a library path, stated as such);
    3. every other row takes the affinity-weighted average of the eigenvectors of its `knn` most similar sampled rows.

`flattened_batch_cluster` is the per-image k-means on that shared embedding (modeling_spectral.py:260-279), on
msvit_kmeans.  With n <= num_sample every row is sampled and the result is the exact dense NCut.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from . import functional as F


def sample_rows(n: int, num_sample: int, seed: int = 0) -> torch.Tensor:
    """Indices of the sampled rows (sorted), from a seeded CPU permutation so that every backend draws the same set."""
    if num_sample >= n:
        return torch.arange(n)
    g = torch.Generator().manual_seed(int(seed))
    return torch.sort(torch.randperm(n, generator=g)[:num_sample]).values


def _pair_affinity(xr: torch.Tensor, xs: torch.Tensor, mode: str, gamma: float, scale: float) -> torch.Tensor:
    """exp(-d / gamma) between two row sets (fp32, cuBLAS GEMM): only used to find and weigh the sampled neighbours."""
    g = xr @ xs.T
    if mode == "cosine":
        d = 1.0 - g * torch.rsqrt((xr * xr).sum(-1).clamp_min(1e-30))[:, None] * torch.rsqrt((xs * xs).sum(-1).clamp_min(1e-30))[None, :]
    elif mode == "rbf":
        d = (0.5 * ((xr * xr).sum(-1)[:, None] + (xs * xs).sum(-1)[None, :]) - g) / scale
    else:
        d = (xr.norm(dim=-1)[:, None] * xs.norm(dim=-1)[None, :] - g) / scale
    return torch.exp(-d.clamp_min(0.0) / gamma)


def _block_iteration(A: torch.Tensor, deg: torch.Tensor, k: int, tol: float, max_iter: int, seed: int):
    """Leading k eigenpairs of D^-1/2 A D^-1/2 for a sample too large for the per-image kernel: orthogonal iteration with
    Rayleigh-Ritz on cuBLAS / cuSOLVER through torch (library path)."""
    n = A.shape[0]
    r = torch.rsqrt(deg)
    m = min(n, k + 8)
    g = torch.Generator(device=A.device).manual_seed(int(seed) + 1)
    U = torch.randn(n, m, generator=g, device=A.device)
    U[:, 0] = torch.sqrt(deg)
    lam = None
    for it in range(max_iter):
        Y = r[:, None] * (A @ (r[:, None] * U))
        Q, _ = torch.linalg.qr(Y)
        H = Q.T @ (r[:, None] * (A @ (r[:, None] * Q)))
        w, W = torch.linalg.eigh(0.5 * (H + H.T))
        U = Q @ W.flip(-1)
        lam_new = w.flip(-1)
        R = r[:, None] * (A @ (r[:, None] * U[:, :k])) - U[:, :k] * lam_new[None, :k]
        if float(R.norm(dim=0).max()) <= tol:
            lam = lam_new
            break
        lam = lam_new
    return U[:, :k].contiguous(), lam[:k].contiguous()


def _sign_fix(V: torch.Tensor) -> torch.Tensor:
    av = V.abs()
    first = (av >= av.max(dim=0, keepdim=True).values).to(torch.int8).argmax(dim=0)
    sgn = torch.sign(V[first, torch.arange(V.shape[1], device=V.device)])
    return V * torch.where(sgn == 0, torch.ones_like(sgn), sgn)[None, :]


@torch.no_grad()
def nystrom_ncut(features: torch.Tensor, num_eig: int, *, num_sample: int = 10000, knn: int = 10, mode: str = "rbf",
                 gamma: float = 3.0, scale: Optional[float] = None, seed: int = 0, tol: float = 2e-5,
                 chunk: int = 8192) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """features [n, D] (CUDA, fp32) -> (eigvecs [n, num_eig], eigvals [num_eig], sampled row ids)."""
    if features.dim() != 2 or not features.is_cuda:
        raise RuntimeError("msvit.nystrom_ncut takes a CUDA [rows, D] matrix; there is no CPU fallback")
    x = features.contiguous().float()
    n, D = x.shape
    s = float(D) if scale is None else float(scale)
    idx = sample_rows(n, num_sample, seed).to(x.device)
    xs = x[idx].contiguous()
    ns = xs.shape[0]
    A, deg = F.affinity(xs[None], mode, gamma, s)                   # tcgen05 Gram + fused distance / exp / degree
    if ns <= 1024 and F.default_block(num_eig) != 0:
        V, lam, _ = F.ncut_eig(A, deg, num_eig, tol=tol)            # per-segment subspace-iteration kernel
        Vs, lam = V[0], lam[0]
    else:
        Vs, lam = _block_iteration(A[0, :, :ns].contiguous(), deg[0], num_eig, tol, 200, seed)
        Vs = _sign_fix(Vs)
    if ns == n:
        return Vs, lam, idx
    out = torch.empty(n, num_eig, dtype=torch.float32, device=x.device)
    out[idx] = Vs
    rest = torch.ones(n, dtype=torch.bool, device=x.device)
    rest[idx] = False
    rest_idx = torch.nonzero(rest).flatten()
    kk = min(knn, ns)
    for r0 in range(0, rest_idx.numel(), chunk):
        rows = rest_idx[r0:r0 + chunk]
        a = _pair_affinity(x[rows], xs, mode, gamma, s)
        w, nb = torch.topk(a, kk, dim=1)
        w = w / w.sum(dim=1, keepdim=True).clamp_min(1e-30)
        out[rows] = (w[:, :, None] * Vs[nb]).sum(dim=1)
    return out, lam, idx


@torch.no_grad()
def flattened_batch_cluster(x: torch.Tensor, ncut_dim: int, n_clusters: int, **kw):
    """The per-image k-means on a batch-level NCut embedding (modeling_spectral.py:253-279): x [B, N, D] ->
    (labels [B, N] int64, eigvecs [B, N, ncut_dim], eigvals [ncut_dim])."""
    B, N, D = x.shape
    V, lam, _ = nystrom_ncut(x.reshape(B * N, D), ncut_dim, **kw)
    V = V.view(B, N, ncut_dim).contiguous()
    labels, n_child, _ = F.kmeans(V, n_clusters)
    return labels, V, lam
