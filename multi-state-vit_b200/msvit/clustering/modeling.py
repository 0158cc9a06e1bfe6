"""Clustering plugin interface -- mirror of the reference's model/clustering/modeling.py:12-36.

Same class names, fields and call contract, so that
`CLUSTERING_CLASSES[config.clustering_config.model_type](config.clustering_config)` in
model/multistate_encoder/modeling_msvitencoder.py:419-423 constructs this implementation unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch
import torch.nn as nn

try:  # the reference derives its configs from transformers' PretrainedConfig (modeling.py:7,12)
    from transformers.configuration_utils import PretrainedConfig as _ConfigBase
except Exception:  # transformers is optional for the hot path itself
    class _ConfigBase:  # type: ignore
        pass


@dataclass
class ClusteringConfig(_ConfigBase):
    model_type: str = None
    ncut_dim: int = None


class ClusteringModule(nn.Module):
    """
    Args:
        parent_indices (`torch.LongTensor` of shape `(batch_size, sequence_length)`):
            Sequence of indices indicating the parent cluster of each token.
        x (`torch.FloatTensor` of shape `(batch_size, sequence_length, hidden_size)`):
            Sequence of hidden-states.

    Returns:
        child_indices (`torch.LongTensor` of shape `(batch_size, sequence_length)`):
            Sequence of indices indicating the child cluster of each token.
    """

    def forward(self, parent_indices: torch.LongTensor, x: torch.FloatTensor, **kwargs: Any) -> torch.LongTensor:
        raise NotImplementedError()
