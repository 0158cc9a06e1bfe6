"""Workload generators of the benchmark (CPU): shapes, determinism per image id, and the structure they promise."""
import torch

from msvit.synthetic import default_scale, hierarchical_tokens, planted_tokens, smooth_tokens
from oracle import ncut_oracle as O


def test_generators_are_per_image_deterministic_and_shardable():
    x, lab = planted_tokens(4, 49, 16, 3)
    x2, lab2 = planted_tokens(2, 49, 16, 3, first=2)
    assert torch.equal(x[2:], x2) and torch.equal(lab[2:], lab2)          # a rank generates exactly its own images
    h, leaf = hierarchical_tokens(3, 256, 16, branch=2, depth=3)
    h2, leaf2 = hierarchical_tokens(1, 256, 16, branch=2, depth=3, first=2)
    assert torch.equal(h[2:], h2) and torch.equal(leaf[2:], leaf2)
    s = smooth_tokens(3, 49, 8)
    assert torch.equal(s[1:2], smooth_tokens(1, 49, 8, first=1))


def test_hierarchical_tokens_have_a_tree_the_oracle_recovers():
    B, N, D = 2, 256, 64
    x, leaf = hierarchical_tokens(B, N, D, branch=2, depth=2)
    assert int(leaf.max()) == 3 and all(int(c) > 0 for c in torch.bincount(leaf[0], minlength=4))
    parent = None
    for level in range(2):
        child, _, lam, _ = O.cluster_tokens(x, parent, ncut_dim=4, n_clusters=2, scale=default_scale(D))
        truth = leaf // (2 ** (1 - level))
        for b in range(B):
            assert torch.equal(O.canonical_relabel(child[b])[0], O.canonical_relabel(truth[b])[0]), (level, b)
        parent = child


def test_smooth_tokens_have_unit_scale_and_no_spectral_gap():
    x = smooth_tokens(1, 196, 256)
    assert 0.9 < float(x.std()) < 1.1
    A = O.affinity(x[0].double(), "rbf", 3.0, default_scale(256))
    _, lam, _ = O.ncut_eig(A, 16)
    ratios = lam[2:16] / lam[1:15]
    assert float(ratios.min()) > 0.4        # the eigenvalues decay gradually: no gap a subspace solver could lean on
