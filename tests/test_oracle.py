"""CPU tests: the oracle against the committed golden vectors and against its own invariants."""
import os

import numpy as np
import pytest
import torch

from oracle import ncut_oracle as O
from msvit.synthetic import default_scale, planted_image, planted_tokens


def test_smoke_8x2_matches_golden(golden_dir):
    # reference's own smoke shape and seed: sandbox/ncut_euclidean.py:13-21
    g = np.load(os.path.join(golden_dir, "smoke_8x2.npz"))
    torch.manual_seed(1212)
    M = torch.randn((8, 2))
    assert np.array_equal(M.numpy(), g["M"])
    A = O.affinity(M.double(), "rbf", 3.0, scale=1.0)
    V, lam, deg = O.ncut_eig(A, 2)
    np.testing.assert_allclose(A.numpy(), g["A"], rtol=1e-12)
    np.testing.assert_allclose(lam.numpy(), g["lam"], rtol=1e-10)
    np.testing.assert_allclose(V.numpy(), g["V"], atol=1e-9)
    assert abs(lam[0].item() - 1.0) < 1e-12  # trivial eigenpair of the normalised affinity
    labels, _, C = O.kmeans(V[:, :2], 2, weight=deg)
    assert np.array_equal(labels.numpy(), g["labels"])


def test_rbf_on_unit_vectors_equals_cosine():
    # the author's check at sandbox/ncut_euclidean.py:23-29
    torch.manual_seed(1212)
    M = torch.nn.functional.normalize(torch.randn(32, 5).double(), dim=-1)
    assert torch.allclose(O.affinity(M, "rbf", 3.0, 1.0), O.affinity(M, "cosine", 3.0), atol=1e-12)


def test_closed_form_matches_reference_restatement():
    # sandbox/test.py:105-118 verbatim math (normprod distance, L = I - D^-1/2 A D^-1/2, eigh, leading k)
    torch.manual_seed(3)
    X = torch.randn(40, 16).double()
    gamma = 1.0
    nX = torch.nn.functional.normalize(X, dim=-1)
    nA = 1.0 - nX @ nX.mT
    A = (X.norm(dim=-1)[:, None] * X.norm(dim=-1)[None, :]) * nA
    A = torch.exp(-A / gamma)
    Dg = A.sum(dim=-1)
    L = torch.eye(len(Dg), dtype=torch.float64) - A * ((Dg[:, None] * Dg[None, :]) ** -0.5)
    E, Vref = torch.linalg.eigh(L)
    ours = O.affinity(X, "normprod", gamma, scale=1.0)
    assert torch.allclose(ours, A, atol=1e-10)
    V, lam, deg = O.ncut_eig(ours, 10)
    assert torch.allclose(1.0 - lam, E[:10], atol=1e-10)
    assert O.subspace_distance(V[:, :10], Vref[:, :10]) < 1e-6


@pytest.mark.parametrize("name", ["c2_196x768", "c2_196x768_b5", "c3_576x1024", "c4_1024x768"])
def test_config_cases_match_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    N, D, K, k, b = int(g["N"]), int(g["D"]), int(g["K"]), int(g["k"]), int(g["b"])
    x, planted = planted_image(b, N, D, K)
    assert np.array_equal(planted.numpy(), g["planted"])
    A = O.affinity(x.double(), "rbf", float(g["gamma"]), float(g["scale"]))
    np.testing.assert_allclose(float(A.sum()), float(g["A_sum"]), rtol=1e-10)
    np.testing.assert_allclose(A[:4, :8].numpy(), g["A_sample"], rtol=1e-10)
    V, lam, deg = O.ncut_eig(A, k + 4)
    np.testing.assert_allclose(lam.numpy(), g["lam"], rtol=1e-8, atol=1e-12)
    labels, _, C = O.kmeans(V[:, :K], K, weight=deg)
    assert C == int(g["n_child"])
    assert np.array_equal(labels.numpy(), g["labels"])
    # planted partition recovered exactly (labels equal modulo permutation)
    pl, _ = O.canonical_relabel(planted)
    assert torch.equal(pl, labels)
    # fp32 oracle agrees with the fp64 golden within the stated tolerances
    A32 = O.affinity(x, "rbf", float(g["gamma"]), float(g["scale"]))
    V32, lam32, deg32 = O.ncut_eig(A32, k)
    np.testing.assert_allclose(lam32.numpy(), g["lam"][:k], rtol=1e-3)
    np.testing.assert_allclose(deg32.numpy(), g["deg"], rtol=1e-4)
    l32, _, _ = O.kmeans(V32[:, :K], K, weight=deg32)
    assert np.array_equal(l32.numpy(), g["labels"])


def test_sign_fix_and_relabel_are_canonical():
    V = torch.tensor([[0.1, -0.9], [-0.8, 0.2], [0.8, 0.9]])
    F = O.sign_fix(V)
    assert F[1, 0] > 0 and F[0, 1] > 0  # first maximal |entry| made positive (ties -> lowest row)
    lab, order = O.canonical_relabel(torch.tensor([2, 2, 0, 1, 0]))
    assert lab.tolist() == [0, 0, 1, 2, 1] and order.tolist() == [2, 0, 1]


def test_kmeans_edge_cases():
    P = torch.tensor([[0.0, 0.0], [0.0, 0.1], [5.0, 5.0], [5.0, 5.1]])
    lab, cen, C = O.kmeans(P, 2)
    assert lab.tolist() == [0, 0, 1, 1] and C == 2
    lab, cen, C = O.kmeans(P, 10)  # more clusters than points -> every point its own cluster
    assert C == 4 and sorted(lab.tolist()) == [0, 1, 2, 3]
    lab, cen, C = O.kmeans(P[:1], 3)
    assert lab.tolist() == [0] and C == 1
    lab, cen, C = O.kmeans(torch.zeros(0, 2), 3)
    assert C == 0 and lab.numel() == 0
    # duplicate points: a seeded centre that attracts nothing keeps its place, ids stay contiguous
    Q = torch.zeros(6, 2)
    lab, cen, C = O.kmeans(Q, 3)
    assert C == 1 and lab.tolist() == [0] * 6


def test_pool_matches_masked_mean():
    torch.manual_seed(0)
    x = torch.randn(2, 9, 4)
    lab = torch.tensor([[0, 1, 1, 2, 0, 5, -1, 2, 2], [3, 3, 3, 3, 3, 3, 3, 3, 3]])
    pooled, counts = O.pool(x, lab, 4)
    assert counts.tolist() == [[2, 2, 3, 0], [0, 0, 0, 9]]
    assert torch.allclose(pooled[0, 2], x[0][[3, 7, 8]].mean(0))
    assert torch.all(pooled[0, 3] == 0) and torch.allclose(pooled[1, 3], x[1].mean(0))


def test_hierarchical_label_contract():
    # msvitencoder.py:491-499: contiguous ids, children of parent p in a contiguous range ordered by p
    B, N, D, K = 3, 96, 32, 4
    x, planted = planted_tokens(B, N, D, K)
    lvl0, _, _, nc0 = O.cluster_tokens(x, None, ncut_dim=4, n_clusters=2, scale=default_scale(D))
    lvl1, _, _, nc1 = O.cluster_tokens(x, lvl0, ncut_dim=4, n_clusters=2, scale=default_scale(D))
    for b in range(B):
        ids = lvl1[b]
        C = int(ids.max()) + 1
        assert sorted(set(ids.tolist())) == list(range(C))
        cum = torch.cumsum(nc1[b], 0)
        parent_of_child = torch.searchsorted(cum, torch.arange(C), side="right")
        assert torch.equal(parent_of_child[ids], lvl0[b])
    # eigenvalue-threshold mode: no eigenvalue above threshold -> exactly one child per parent
    one, _, _, nc = O.cluster_tokens(x, lvl0, ncut_dim=4, eigenvalue_threshold=2.0, scale=default_scale(D))
    assert torch.equal(one, lvl0) and torch.all(nc[nc > 0] == 1)


def test_global_kmeans_recovers_planted_centres():
    from msvit.synthetic import planted_features
    f = planted_features(4000, 16, 8, noise=0.05)
    g = torch.Generator().manual_seed(1212)
    centres = torch.randn(8, 16, generator=g)
    C, labels, counts = O.global_kmeans(f, 8, 10, init=centres + 0.01)
    assert int(counts.sum()) == 4000
    assert torch.allclose(C, centres, atol=0.02)
    # shard-sum equivalence: sums/counts over two halves add up (what the NCCL allreduce relies on)
    C2, _, _ = O.global_kmeans(torch.cat([f[2000:], f[:2000]]), 8, 10, init=centres + 0.01)
    assert torch.allclose(C, C2, atol=1e-5)


def test_synthetic_generators_are_shard_consistent():
    from msvit.synthetic import planted_features
    x, _ = planted_tokens(4, 12, 8, 3)
    x2, _ = planted_tokens(2, 12, 8, 3, first=2)
    assert torch.equal(x[2:], x2)
    f = planted_features(300, 8, 5, chunk=128)
    f2 = planted_features(100, 8, 5, first=150, chunk=128)
    assert torch.equal(f[150:250], f2)


def test_attention_mask_structure():
    # msvitencoder.py:426-467: hand-checked tiny case, 2 clusters in image 0 and 1 cluster in image 1 (C = 2)
    lab = torch.tensor([[0, 1, 0], [0, 0, 0]])
    m = O.attention_mask(lab)[:, 0]
    assert m.shape == (2, 7, 7)
    T0, R0, T1, R1, a, b, c = range(7)
    img = m[0]
    assert img[a, c] and img[c, a] and not img[a, b]            # tokens a, c share cluster 0
    assert img[T0, a] and img[T0, c] and not img[T0, b] and img[T1, b]
    assert img[a, R0] and img[b, R1] and not img[b, R0]
    assert img[R0, T0] and img[R0, T1] and img[R1, T0] and not img[T0, R0]
    one = m[1]
    assert one[R0, T0] and not one[R1, T0] and not one[R0, T1]   # image 1 has a single live cluster
    assert not one[T1].any() and one[T0, a] and one[a, R0]


def test_kway_ncut_recovers_a_planted_partition_and_is_rotation_invariant():
    # axis-aligned discretisation (Yu & Shi 2003; ncut_pytorch.kway_ncut at modeling_spectral.py:136-138): on a clean
    # planted mixture it recovers the partition, and rotating the embedding columns does not change the labels
    from msvit.synthetic import default_scale, planted_image
    x, lab = planted_image(3, 120, 64, 5)
    A = O.affinity(x.double(), "rbf", 3.0, default_scale(64))
    V, lam, deg = O.ncut_eig(A, 5)
    labels, C, R = O.kway_ncut(V, 5, weight=deg)
    assert C == 5 and torch.equal(labels, O.canonical_relabel(lab)[0])
    assert torch.allclose(R.T @ R, torch.eye(5, dtype=R.dtype), atol=1e-9)          # a rotation
    Q, _ = torch.linalg.qr(torch.randn(5, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(0)))
    labels_rot, C_rot, _ = O.kway_ncut(V @ Q, 5, weight=deg)
    assert C_rot == 5 and torch.equal(labels_rot, labels)
    # same partition as k-means on this embedding
    km, _, _ = O.kmeans(V[:, :5], 5, weight=deg)
    assert torch.equal(km, labels)
