"""Spectral (NCut) clustering plugin -- mirror of model/clustering/modeling_spectral.py:42-94.

Same config fields (`ncut_dim`, `ncut_dist`, `eigenvalue_threshold`, `cluster_size_threshold`) and the same
`forward(parent_indices, x) -> child_indices` contract; the body is the B200 kernel sequence in
msvit.functional.cluster_tokens instead of ncut_pytorch + cuML calls in a Python loop.

Differences from the reference that are deliberate:
  * segments are (image, parent) -- the per-image variant at modeling_spectral.py:260-279 -- not parents
    pooled over the whole batch (:83-84);
  * a parent with no eigenvalue above the threshold yields exactly one child AND advances the child offset
    (the reference adds 0 at :94, which would make two parents share an id);
  * exact (tolerance-driven) eigenvectors and deterministic k-means seeding instead of randomised ones;
  * `cluster_size_threshold` is accepted and ignored: the reference only reads it in its plotting branch
    (modeling_spectral.py:110-113).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Literal, Optional

import torch

from .. import functional as F
from .modeling import ClusteringConfig, ClusteringModule


@dataclass
class SpectralClusteringConfig(ClusteringConfig):
    model_type: str = "spectral"
    ncut_dist: Literal["rbf", "cosine"] = None
    eigenvalue_threshold: float = None
    cluster_size_threshold: float = None
    # ---- additions.  affinity_focal_gamma = 3.0 is the reference's NCUT(...) argument (modeling_spectral.py:59).
    # distance_scale is OURS: the rbf distance is 1/2 |xi - xj|^2 / s with s = hidden size by default, so that gamma does
    # not depend on the width of the model; s = 1.0 gives the raw form of sandbox/ncut_euclidean.py:19,23-29.  ncut-pytorch's
    # own feature scaling cannot be inspected here (the package is absent), so an eigenvalue_threshold tuned against
    # the reference's eigenvalues has to be re-tuned.  "normprod" (sandbox/test.py:108-110) is not a positive
    # semi-definite kernel in general: the solver returns the pairs largest in magnitude.
    affinity_focal_gamma: float = 3.0
    distance_scale: Optional[float] = None   # None -> hidden size (rbf / normprod)
    n_clusters: Optional[int] = None         # fixed children per parent instead of the eigenvalue threshold
    discretise: Literal["kmeans", "axis_align"] = "kmeans"   # axis_align = kway_ncut (modeling_spectral.py:136-138)
    kmeans_iters: int = 100
    eig_iters: int = 60
    eig_tol: float = 2e-5


class SpectralClustering(ClusteringModule):
    def __init__(self, config: SpectralClusteringConfig):
        super().__init__()
        self.config = config
        self.last_output: Optional[F.ClusterOutput] = None   # eigenpairs, iterations and the converged flags of the last call

    def cluster(self, parent_indices: Optional[torch.LongTensor], x: torch.Tensor, **kwargs: Any) -> F.ClusterOutput:
        c = self.config
        thr = c.eigenvalue_threshold
        if c.n_clusters is None and thr is None:
            raise ValueError("SpectralClusteringConfig needs eigenvalue_threshold or n_clusters")
        return F.cluster_tokens(
            x, parent_indices, ncut_dim=c.ncut_dim, n_clusters=c.n_clusters, eigenvalue_threshold=thr,
            mode=c.ncut_dist or "rbf", gamma=c.affinity_focal_gamma, scale=c.distance_scale,
            kmeans_iters=c.kmeans_iters, eig_iters=c.eig_iters, eig_tol=c.eig_tol, discretise=c.discretise,
            n_parents=kwargs.get("n_parents"), want_pool=kwargs.get("want_pool", False),
            pool_k=kwargs.get("pool_k"))

    @torch.no_grad()
    def forward(self, parent_indices: torch.LongTensor, x: torch.FloatTensor, **kwargs: Any) -> torch.LongTensor:
        bsz, N = parent_indices.shape
        if x.shape[:2] != (bsz, N):
            raise ValueError("parent_indices and x disagree on (batch, tokens)")
        self.last_output = self.cluster(parent_indices, x, **kwargs)
        return self.last_output.labels
