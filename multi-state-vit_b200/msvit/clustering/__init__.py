"""Registry of clustering plugins -- mirror of model/clustering/__init__.py:7-10.

The reference also registers "fps" (FPSClustering); that class cannot be constructed or run in the reference
itself (modeling_fps.py:25,37,40) and is outside the hot path, so only
"spectral" is provided.
"""
from typing import Dict

from .modeling import ClusteringConfig, ClusteringModule
from .modeling_spectral import SpectralClustering, SpectralClusteringConfig

CLUSTERING_CLASSES: Dict[str, type] = {
    "spectral": SpectralClustering,
}

__all__ = ["ClusteringConfig", "ClusteringModule", "SpectralClustering", "SpectralClusteringConfig",
           "CLUSTERING_CLASSES"]
