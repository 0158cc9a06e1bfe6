"""msvit -- B200 (sm_100a) implementation of multi-state-ViT's token-grouping hot path.

    from msvit import CLUSTERING_CLASSES, SpectralClusteringConfig     # the reference's plugin interface
    from msvit import cluster_tokens, pool                             # functional surface

All hot-path compute runs in libmsvit.so (csrc/, C ABI in include/msvit.h); there is no CPU fallback.  Two regimes
outside the headline path call cuSOLVER / cuBLAS through torch and say so in their docstrings: ncut_dim > 32
(functional.dense_ncut_eig) and Nystrom samples of more than 1024 rows (nystrom._block_iteration).
"""
from . import _lib
from .clustering import (CLUSTERING_CLASSES, ClusteringConfig, ClusteringModule, SpectralClustering,
                         SpectralClusteringConfig)
from .global_kmeans import GlobalKMeansPlan, GlobalKMeansResult, global_kmeans
from .nystrom import flattened_batch_cluster, nystrom_ncut
from .functional import ClusterOutput, ClusterPlan, HostClusterer, HostResult, affinity, attention_mask, cluster_attention_stats, cluster_tokens, kmeans, ncut_eig, pool

__all__ = ["CLUSTERING_CLASSES", "ClusteringConfig", "ClusteringModule", "SpectralClustering",
           "SpectralClusteringConfig", "ClusterOutput", "ClusterPlan", "HostClusterer", "HostResult", "affinity", "attention_mask", "cluster_attention_stats", "cluster_tokens", "kmeans", "ncut_eig", "pool",
           "GlobalKMeansPlan", "GlobalKMeansResult", "global_kmeans", "nystrom_ncut", "flattened_batch_cluster", "_lib"]
