"""msvit -- B200 (sm_100a) implementation of multi-state-ViT's token-grouping hot path.

    from msvit import CLUSTERING_CLASSES, SpectralClusteringConfig     # the reference's plugin interface
    from msvit import cluster_tokens, pool                             # functional surface

All compute runs in libmsvit.so (csrc/, C ABI in include/msvit.h); there is no CPU or library fallback.
"""
from . import _lib
from .clustering import (CLUSTERING_CLASSES, ClusteringConfig, ClusteringModule, SpectralClustering,
                         SpectralClusteringConfig)
from .global_kmeans import GlobalKMeansPlan, GlobalKMeansResult, global_kmeans
from .functional import ClusterOutput, ClusterPlan, HostClusterer, HostResult, affinity, attention_mask, cluster_attention_stats, cluster_tokens, kmeans, ncut_eig, pool

__all__ = ["CLUSTERING_CLASSES", "ClusteringConfig", "ClusteringModule", "SpectralClustering",
           "SpectralClusteringConfig", "ClusterOutput", "ClusterPlan", "HostClusterer", "HostResult", "affinity", "attention_mask", "cluster_attention_stats", "cluster_tokens", "kmeans", "ncut_eig", "pool",
           "GlobalKMeansPlan", "GlobalKMeansResult", "global_kmeans", "_lib"]
