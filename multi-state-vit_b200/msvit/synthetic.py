"""Synthetic ViT token tensors for tests and benchmarks.

iid Gaussian tokens give a degenerate NCut spectrum (lambda ~ [1, .004, .004, ...]) whose
eigenvectors and labels are numerically arbitrary, so every workload here is a planted
mixture: per image, K random centres with unequal class sizes plus isotropic noise.
Generation is on the CPU with a per-image seed so that every rank / shard can produce
exactly its own images without communication.
"""
from __future__ import annotations

from typing import Tuple

import torch

# name -> (B, N, D, K)   (BASELINE.json configs[0..3])
CONFIGS = {
    "C1": (8, 196, 768, 8),
    "C2": (1024, 196, 768, 8),
    "C3": (512, 576, 1024, 16),
    "C4": (256, 1024, 768, 4),
}


def default_scale(D: int) -> float:
    """Distance scale used by the benchmark workloads: s = D / 4 (about median |xi-xj|^2 / 10)."""
    return D / 4.0


def planted_image(b: int, N: int, D: int, K: int, noise: float = 0.5, seed: int = 1212) -> Tuple[torch.Tensor, torch.Tensor]:
    """Image `b` of a workload: (x [N, D] fp32, planted labels [N] int64)."""
    g = torch.Generator().manual_seed(seed + b)
    centres = torch.randn(K, D, generator=g)
    # class sizes proportional to c + 1 (unequal sizes => distinct eigenvalues), exact rather than sampled so
    # that no class is ever missing: a missing class would put a wanted eigenvector into the noise bulk,
    # where eigenvectors -- and hence labels -- are numerically arbitrary for ANY solver
    w = torch.arange(1, K + 1, dtype=torch.float64)
    sizes = torch.floor(w / w.sum() * N).long()
    sizes[K - 1] += N - int(sizes.sum())
    lab = torch.repeat_interleave(torch.arange(K), sizes)[torch.randperm(N, generator=g)]
    x = centres[lab] + noise * torch.randn(N, D, generator=g)
    return x, lab


def planted_tokens(B: int, N: int, D: int, K: int, noise: float = 0.5, seed: int = 1212, first: int = 0,
                   out: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Images first .. first+B-1 -> (x [B, N, D] fp32, labels [B, N])."""
    x = out if out is not None else torch.empty(B, N, D)
    lab = torch.empty(B, N, dtype=torch.long)
    for i in range(B):
        xi, li = planted_image(first + i, N, D, K, noise, seed)
        x[i].copy_(xi)
        lab[i] = li
    return x, lab


def hierarchical_image(b: int, N: int, D: int, branch: int = 4, depth: int = 3, spread=(1.0, 0.7, 0.5),
                       noise: float = 0.3, seed: int = 1212) -> Tuple[torch.Tensor, torch.Tensor]:
    """Image `b` of the hierarchical workload (C4): a planted tree of `depth` levels with `branch` children per node.
    A node's centre is its parent's centre + spread[level] * randn(D); a token is its leaf's centre + noise * randn(D).
    Children of a node have unequal sizes (proportional to branch + c), exact rather than sampled.  Re-clustering a
    level-l cluster therefore has `branch` sub-clusters to find, which a flat mixture does not offer.
    Returns (x [N, D] fp32, leaf ids [N] int64 in [0, branch**depth))."""
    g = torch.Generator().manual_seed(seed + b)
    centres = torch.zeros(1, D)
    sizes = torch.tensor([N])
    w = torch.arange(branch, 2 * branch, dtype=torch.float64)
    for level in range(depth):
        centres = (centres[:, None, :] + spread[level] * torch.randn(centres.shape[0], branch, D, generator=g)).reshape(-1, D)
        new = []
        for n_node in sizes.tolist():
            sz = torch.floor(w / w.sum() * n_node).long()
            sz[branch - 1] += n_node - int(sz.sum())
            new.append(sz)
        sizes = torch.cat(new)
    leaf = torch.repeat_interleave(torch.arange(centres.shape[0]), sizes)[torch.randperm(N, generator=g)]
    x = centres[leaf] + noise * torch.randn(N, D, generator=g)
    return x, leaf


def hierarchical_tokens(B: int, N: int, D: int, branch: int = 4, depth: int = 3, seed: int = 1212, first: int = 0
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Images first .. first+B-1 of the hierarchical workload -> (x [B, N, D] fp32, leaf ids [B, N])."""
    x = torch.empty(B, N, D)
    leaf = torch.empty(B, N, dtype=torch.long)
    for i in range(B):
        x[i], leaf[i] = hierarchical_image(first + i, N, D, branch, depth, seed=seed)
    return x, leaf


def smooth_tokens(B: int, N: int, D: int, n_freq: int = 64, alpha: float = 1.5, noise: float = 0.1, seed: int = 1212,
                  first: int = 0) -> torch.Tensor:
    """Tokens WITHOUT planted clusters: every feature channel is a random smooth field over the patch grid (2-D cosine
    modes with power-law amplitudes (1 + u^2 + v^2)^(-alpha/2), unit variance per channel) plus white noise.  The NCut
    spectrum of such an image decays gradually instead of showing a gap after K eigenvalues -- the hard case for a
    subspace solver, used by bench.py to report iteration counts on non-planted data.  x [B, N, D] fp32."""
    g = int(round(N ** 0.5))
    if g * g == N:
        ii = (torch.arange(g, dtype=torch.float64) + 0.5) / g
        modes = [(u, v) for u in range(g) for v in range(g)]
        modes.sort(key=lambda m: (m[0] ** 2 + m[1] ** 2, m))
        modes = modes[1:n_freq + 1]                      # drop the constant mode
        basis = torch.stack([(torch.cos(torch.pi * u * ii)[:, None] * torch.cos(torch.pi * v * ii)[None, :]).reshape(N)
                             for u, v in modes], 1)      # [N, F]
        amp = torch.tensor([(1.0 + u * u + v * v) ** (-alpha / 2) for u, v in modes], dtype=torch.float64)
    else:
        ii = (torch.arange(N, dtype=torch.float64) + 0.5) / N
        freqs = torch.arange(1, n_freq + 1, dtype=torch.float64)
        basis = torch.cos(torch.pi * freqs[None, :] * ii[:, None])
        amp = (1.0 + freqs ** 2) ** (-alpha / 2)
    basis = basis * amp[None, :]
    basis = (basis / basis.pow(2).sum(1).mean().sqrt()).float()   # unit average variance per channel
    x = torch.empty(B, N, D)
    for i in range(B):
        gen = torch.Generator().manual_seed(seed + 7919 * (first + i))
        coef = torch.randn(basis.shape[1], D, generator=gen)
        x[i] = basis @ coef + noise * torch.randn(N, D, generator=gen)
    return x


def planted_features(n: int, D: int, k: int, noise: float = 0.5, seed: int = 1212, first: int = 0,
                     chunk: int = 65536) -> torch.Tensor:
    """Rows first .. first+n-1 of the dataset-level (DeepCluster-style) workload: centres[lab] + noise."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(k, D, generator=g)
    out = torch.empty(n, D)
    start = (first // chunk) * chunk
    pos = 0
    c = start
    while pos < n:
        gc = torch.Generator().manual_seed(seed + 1 + c // chunk)
        lab = torch.randint(0, k, (chunk,), generator=gc)
        blk = centres[lab] + noise * torch.randn(chunk, D, generator=gc)
        lo = max(first, c) - c
        hi = min(first + n, c + chunk) - c
        out[pos:pos + hi - lo] = blk[lo:hi]
        pos += hi - lo
        c += chunk
    return out
