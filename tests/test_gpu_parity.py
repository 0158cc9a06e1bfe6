"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI via the torch
custom ops, against the CPU oracle on the same seeded inputs and against the committed golden vectors.

Bars (BASELINE.json north_star / SURVEY.md section 8d):
  labels            identical modulo permutation  -> exact equality after first-occurrence relabelling
  affinity, eigenvalues, pooled tokens            -> rtol 1e-3 (fp32 accumulation)
  eigenvectors      per column, up to the gap: |v - v_ref| <= 1e-3 + 2e-4 / gap, sign canonical
The oracle is fed the operand the tensor core actually reads: bf16-rounded tokens for bf16 input,
TF32-truncated tokens for fp32 input.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ncut_oracle as O  # noqa: E402
import msvit  # noqa: E402
from msvit import functional as F, ops  # noqa: E402
from msvit.synthetic import default_scale, planted_image, planted_tokens  # noqa: E402

DEV = "cuda:0"
RTOL = 1e-3


def as_seen_by_tensor_core(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return O.round_to_bf16(x) if dtype == torch.bfloat16 else O.round_to_tf32(x)


def canon(labels: torch.Tensor) -> torch.Tensor:
    return O.canonical_relabel(labels.cpu())[0]


def check_eigvecs(V, lam_all, Vref, k):
    """Column-wise comparison with a gap-aware tolerance; lam_all holds >= k+1 reference eigenvalues."""
    V = V.double().cpu()
    Vref = Vref.double()
    for j in range(k):
        gaps = [abs(lam_all[j] - lam_all[i]) for i in range(len(lam_all)) if i != j]
        gap = float(min(gaps))
        err = float(torch.linalg.norm(V[:, j] - Vref[:, j]))
        assert err <= 1e-3 + 2e-4 / max(gap, 1e-9), f"eigvec {j}: err {err:.2e}, gap {gap:.2e}"
    assert O.subspace_distance(V[:, :k], Vref[:, :k]) < 5e-3


# ----------------------------------------------------------------------------------------- pooling
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 196, 768, 8), (2, 50, 20, 3), (2, 33, 7, 5), (1, 1024, 768, 12)])
def test_pool_matches_oracle(dtype, shape):
    B, N, D, K = shape
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, N, D, generator=g).to(dtype)
    lab = torch.randint(-1, K + 1, (B, N), generator=g)  # includes ids outside [0, K): ignored
    pooled, counts = msvit.pool(x.to(DEV), lab.to(DEV), K)
    rp, rc = O.pool(x.double(), lab, K)
    assert torch.equal(counts.cpu(), rc)
    torch.testing.assert_close(pooled.cpu().double(), rp, rtol=1e-5, atol=1e-5)


def test_pool_empty_cluster_and_single_token():
    x = torch.arange(2 * 4 * 8, dtype=torch.float32).view(2, 4, 8)
    lab = torch.tensor([[0, 0, 0, 0], [2, 2, 2, 1]])
    pooled, counts = msvit.pool(x.to(DEV), lab.to(DEV), 3)
    assert counts.cpu().tolist() == [[4, 0, 0], [0, 1, 3]]
    assert torch.all(pooled[0, 1:] == 0)
    assert torch.equal(pooled[1, 1].cpu(), x[1, 3])
    assert torch.allclose(pooled[0, 0].cpu(), x[0].mean(0))


# ----------------------------------------------------------------------------------------- k-means
def test_kmeans_matches_oracle_on_spectral_embedding():
    B, N, D, K = 8, 196, 768, 8
    x, _ = planted_tokens(B, N, D, K)
    Vs, degs, labs = [], [], []
    for b in range(B):
        V, lam, deg = O.ncut_eig(O.affinity(x[b], "rbf", 3.0, default_scale(D)), K)
        l, _, C = O.kmeans(V, K, weight=deg)
        Vs.append(V); degs.append(deg); labs.append(l)
    V = torch.stack(Vs).to(DEV)
    deg = torch.stack(degs).to(DEV)
    labels, n_child, centres = F.kmeans(V, K, weight=deg)
    assert torch.equal(labels.cpu(), torch.stack(labs))
    assert n_child.cpu().tolist() == [K] * B


def test_kmeans_edge_cases():
    # fewer points than clusters, duplicate points, eigenvalue-threshold selection of K, caller-supplied seeds
    P = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.1, 0.0], [5.0, 5.0, 0.0], [5.0, 5.1, 0.0]])
    V = P[None].to(DEV)
    labels, n_child, _ = F.kmeans(V, 2)
    assert labels.cpu().tolist() == [[0, 0, 1, 1]] and n_child.item() == 2
    labels, n_child, _ = F.kmeans(V, 3)  # K = min(K, n, columns) = 3
    ref, _, C = O.kmeans(P, 3)
    assert labels.cpu()[0].tolist() == ref.tolist() and n_child.item() == C
    Z = torch.zeros(1, 6, 3, device=DEV)
    labels, n_child, _ = F.kmeans(Z, 3)
    assert labels.cpu().tolist() == [[0] * 6] and n_child.item() == 1
    lam = torch.tensor([[1.0, 0.5, 0.01]], device=DEV)
    labels, n_child, _ = F.kmeans(V, None, eigvals=lam, eigenvalue_threshold=0.1)
    ref, _, C = O.kmeans(P[:, :2], 2)
    assert labels.cpu()[0].tolist() == ref.tolist() and n_child.item() == 2
    labels, n_child, _ = F.kmeans(V, None, eigvals=lam, eigenvalue_threshold=5.0)  # none above -> one child
    assert labels.cpu().tolist() == [[0, 0, 0, 0]] and n_child.item() == 1
    init = torch.tensor([[[5.0, 5.0], [0.0, 0.0]]], device=DEV)
    labels, n_child, _ = F.kmeans(V[:, :, :2].contiguous(), 2, init=init)
    assert labels.cpu().tolist() == [[0, 0, 1, 1]]  # canonical ids do not depend on the seed order


# ----------------------------------------------------------------------------------------- eigensolver
@pytest.mark.parametrize("case", [(196, 768, 8, 8), (64, 32, 3, 4), (10, 16, 2, 8), (576, 1024, 16, 16),
                                  (784, 768, 12, 100)])   # the last one: the author's shape (sandbox/test.py:22,47-52,66)
def test_ncut_eig_matches_exact_eigh(case):
    N, D, Kp, k = case
    B = 3
    x, _ = planted_tokens(B, N, D, Kp)
    lda = ops.lda_of(N)
    A = torch.zeros(B, N, lda)
    refs = []
    for b in range(B):
        Ab = O.affinity(x[b], "rbf", 3.0, default_scale(D))
        A[b, :, :N] = Ab
        refs.append(O.ncut_eig(Ab.double(), min(k + 6, N)))
    deg = A.sum(-1)
    V, lam, iters = F.ncut_eig(A.to(DEV), deg.to(DEV), k)
    assert int(iters.max()) < 60, f"eigensolver hit the iteration cap: {iters.cpu().tolist()}"
    for b in range(B):
        Vref, lref, _ = refs[b]
        kk = min(k, N)
        np.testing.assert_allclose(lam[b, :kk].cpu().numpy(), lref[:kk].numpy(), rtol=RTOL, atol=1e-6)
        check_eigvecs(V[b], lref.tolist(), Vref, min(kk, Kp))
        # canonical sign: the largest-|entry| of every wanted column is positive
        Vb = V[b].cpu()
        for j in range(min(kk, Kp)):
            assert Vb[torch.argmax(Vb[:, j].abs()), j] > 0
        if N > k:
            assert torch.all(lam[b, :-1] >= lam[b, 1:] - 1e-6)  # descending


# ----------------------------------------------------------------------------------------- affinity
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(4, 196, 768), (2, 64, 64), (3, 100, 40), (2, 300, 128), (1, 576, 1024),
                                   (1, 1024, 768), (2, 37, 24)])
def test_affinity_degree_matches_oracle(dtype, shape):
    B, N, D = shape
    x, _ = planted_tokens(B, N, D, 4)
    xq = as_seen_by_tensor_core(x, dtype)
    A, deg = F.affinity(x.to(dtype).to(DEV), "rbf", 3.0, default_scale(D))
    assert A.shape == (B, N, ops.lda_of(N))
    if ops.lda_of(N) > N:
        assert torch.all(A[:, :, N:] == 0)
    for b in range(B):
        ref = O.affinity(xq[b].double(), "rbf", 3.0, default_scale(D))
        torch.testing.assert_close(A[b, :, :N].cpu().double(), ref, rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(deg[b].cpu().double(), ref.sum(-1), rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("mode", ["cosine", "normprod"])
def test_affinity_other_distances(mode):
    B, N, D = 2, 196, 768
    x, _ = planted_tokens(B, N, D, 4)
    xq = O.round_to_bf16(x)
    A, deg = F.affinity(x.bfloat16().to(DEV), mode, 3.0, default_scale(D))
    for b in range(B):
        ref = O.affinity(xq[b].double(), mode, 3.0, default_scale(D))
        torch.testing.assert_close(A[b, :, :N].cpu().double(), ref, rtol=RTOL, atol=1e-6)


# ----------------------------------------------------------------------------------------- whole path
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cluster_tokens_c1_matches_oracle(dtype):
    # BASELINE.json configs[0]: ViT-B/16 224px (196 tokens, d=768), batch 8, k=8
    B, N, D, K = 8, 196, 768, 8
    x, planted = planted_tokens(B, N, D, K)
    out = msvit.cluster_tokens(x.to(dtype).to(DEV), ncut_dim=K, n_clusters=K, scale=default_scale(D), keep_affinity=True)
    xq = as_seen_by_tensor_core(x, dtype)
    child, eigvecs, eigvals, n_children = O.cluster_tokens(xq.double(), None, ncut_dim=K, n_clusters=K,
                                                           scale=default_scale(D))
    assert out.labels.dtype == torch.int64 and out.labels.shape == (B, N)
    assert torch.equal(out.labels.cpu(), child)
    np.testing.assert_allclose(out.eigvals[:, 0].cpu().numpy(), eigvals[:, 0].numpy(), rtol=RTOL)
    pooled_ref, counts_ref = O.pool(x.to(dtype).double(), child, K)
    assert torch.equal(out.counts.cpu(), counts_ref)
    torch.testing.assert_close(out.pooled.cpu().double(), pooled_ref, rtol=RTOL, atol=1e-5)
    for b in range(B):
        ref = O.affinity(xq[b].double(), "rbf", 3.0, default_scale(D))
        torch.testing.assert_close(out.affinity[b].cpu().double(), ref, rtol=RTOL, atol=1e-6)
        assert torch.equal(canon(planted[b]), out.labels[b].cpu())  # planted partition recovered


def test_cluster_tokens_threshold_mode_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "c2_196x768.npz"))
    N, D, K = int(g["N"]), int(g["D"]), int(g["K"])
    x, _ = planted_image(0, N, D, K)
    out = msvit.cluster_tokens(x[None].to(DEV), ncut_dim=K, n_clusters=K, scale=float(g["scale"]))
    np.testing.assert_allclose(out.eigvals[0, 0].cpu().numpy(), g["lam"][:K], rtol=RTOL)
    assert np.array_equal(out.labels[0].cpu().numpy(), g["labels"].astype(np.int64))
    np.testing.assert_allclose(out.degree[0].cpu().numpy(), g["deg"], rtol=RTOL)
    np.testing.assert_allclose(out.pooled[0, :, :16].cpu().numpy(), g["pooled_sample"], rtol=RTOL, atol=1e-5)
    # eigenvalue-threshold selection of the number of children (modeling_spectral.py:87)
    thr = msvit.cluster_tokens(x[None].to(DEV), ncut_dim=K, eigenvalue_threshold=0.1, scale=float(g["scale"]))
    child, _, _, nc = O.cluster_tokens(O.round_to_tf32(x[None]).double(), None, ncut_dim=K,
                                       eigenvalue_threshold=0.1, scale=float(g["scale"]))
    assert thr.n_child.cpu().tolist() == nc.tolist()
    assert torch.equal(thr.labels.cpu(), child)


@pytest.mark.parametrize("name", ["c3_576x1024", "c4_1024x768"])
def test_large_token_counts_match_golden(golden_dir, name):
    # BASELINE.json configs[2], configs[3] (one image each): streaming-affinity eigensolver path
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    N, D, K, k = int(g["N"]), int(g["D"]), int(g["K"]), int(g["k"])
    x, _ = planted_image(int(g["b"]), N, D, K)
    out = msvit.cluster_tokens(x[None].to(DEV), ncut_dim=k, n_clusters=K, scale=float(g["scale"]))
    np.testing.assert_allclose(out.eigvals[0, 0, :K].cpu().numpy(), g["lam"][:K], rtol=RTOL)
    assert np.array_equal(out.labels[0].cpu().numpy(), g["labels"].astype(np.int64))
    np.testing.assert_allclose(out.pooled[0, :, :16].cpu().numpy(), g["pooled_sample"], rtol=RTOL, atol=1e-5)
    assert np.array_equal(out.counts[0].cpu().numpy(), g["counts"])


def test_hierarchical_reclustering_matches_oracle():
    # configs[3] in miniature: 3 levels, each parent segment re-clustered (msvitencoder.py:487-509)
    B, N, D = 4, 256, 128
    x, _ = planted_tokens(B, N, D, 8)
    s = default_scale(D)
    parent_gpu = None
    parent_cpu = None
    xq = O.round_to_tf32(x).double()
    for level in range(3):
        out = msvit.cluster_tokens(x.to(DEV), parent_gpu, ncut_dim=4, n_clusters=2, scale=s)
        child, _, _, nc = O.cluster_tokens(xq, parent_cpu, ncut_dim=4, n_clusters=2, scale=s)
        assert torch.equal(out.labels.cpu(), child), f"level {level}"
        assert out.n_child.cpu().tolist() == nc.tolist()
        # caller contract: contiguous ids, parent recoverable by cumsum + searchsorted
        for b in range(B):
            C = int(out.labels[b].max()) + 1
            assert sorted(set(out.labels[b].cpu().tolist())) == list(range(C))
            if parent_cpu is not None:
                cum = torch.cumsum(out.n_child[b].cpu().long(), 0)
                poc = torch.searchsorted(cum, torch.arange(C), side="right")
                assert torch.equal(poc[out.labels[b].cpu()], parent_cpu[b])
        parent_gpu, parent_cpu = out.labels, child


def test_ragged_and_tiny_segments():
    # parents of very different sizes, including single-token and empty parents
    B, N, D = 2, 64, 32
    x, _ = planted_tokens(B, N, D, 3)
    parent = torch.zeros(B, N, dtype=torch.long)
    parent[0, :1] = 1          # one-token parent
    parent[0, 1:4] = 3         # three tokens, parent id 2 is empty
    parent[1, 10:40] = 2
    out = msvit.cluster_tokens(x.to(DEV), parent.to(DEV), ncut_dim=4, n_clusters=3, scale=default_scale(D))
    child, _, _, nc = O.cluster_tokens(O.round_to_tf32(x).double(), parent, ncut_dim=4, n_clusters=3,
                                       scale=default_scale(D))
    assert out.n_child.cpu().tolist() == nc.tolist()
    assert torch.equal(out.labels.cpu(), child)


def test_module_forward_is_a_drop_in():
    # the call the ms-ViT backbone makes: cluster_module(cluster_indices, hidden_states)  (msvitencoder.py:490)
    B, N, D, K = 4, 196, 768, 8
    x, planted = planted_tokens(B, N, D, K)
    cfg = msvit.SpectralClusteringConfig(ncut_dim=K, ncut_dist="rbf", eigenvalue_threshold=0.05,
                                         cluster_size_threshold=0.07, distance_scale=default_scale(D))
    module = msvit.CLUSTERING_CLASSES[cfg.model_type](cfg).to(DEV)
    parents = torch.zeros(B, N, dtype=torch.long, device=DEV)  # msvitencoder.py:478
    child = module(parents, x.to(DEV))
    assert child.dtype == torch.int64 and child.shape == (B, N) and child.device.type == "cuda"
    ref, _, _, _ = O.cluster_tokens(O.round_to_tf32(x).double(), None, ncut_dim=K, eigenvalue_threshold=0.05,
                                    scale=default_scale(D))
    assert torch.equal(child.cpu(), ref)
    n_child = child.max(dim=1).values + 1  # what the caller computes (msvitencoder.py:491)
    assert n_child.cpu().tolist() == [K] * B


def test_full_size_c2_properties():
    # BASELINE.json configs[1] at full size (batch 1024): oracle-free, size-independent properties
    B, N, D, K = 1024, 196, 768, 8
    x, planted = planted_tokens(B, N, D, K)
    xg = x.to(DEV)
    out = msvit.cluster_tokens(xg, ncut_dim=K, n_clusters=K, scale=default_scale(D))
    labels = out.labels.cpu()
    for b in range(B):
        assert torch.equal(canon(planted[b]), labels[b]), f"image {b}: planted partition not recovered"
    assert torch.allclose(out.eigvals[:, 0, 0].cpu(), torch.ones(B), atol=1e-4)       # trivial eigenvalue
    assert int(out.counts.sum()) == B * N                                              # every token pooled once
    recon = (out.pooled * out.counts[..., None]).sum(1) / N                            # count-weighted mean of means
    torch.testing.assert_close(recon.cpu(), x.mean(1), rtol=1e-3, atol=1e-4)
    again = msvit.cluster_tokens(xg, ncut_dim=K, n_clusters=K, scale=default_scale(D))
    assert torch.equal(again.labels, out.labels) and torch.equal(again.pooled, out.pooled)  # bit-reproducible
    # batch sharding equivalence: a shard clustered alone gives the same labels (no cross-image coupling)
    part = msvit.cluster_tokens(xg[512:640], ncut_dim=K, n_clusters=K, scale=default_scale(D))
    assert torch.equal(part.labels, out.labels[512:640])


# ----------------------------------------------------------------------------------------- next row
@pytest.mark.parametrize("shape", [(4, 196, 8), (3, 50, 5), (2, 33, 3), (1, 1024, 12), (5, 7, 7)])
def test_attention_mask_matches_reference_restatement(shape):
    # MultiStateViTEncoderBackbone._construct_attention_mask (msvitencoder.py:426-467): bit-exact boolean mask
    B, N, K = shape
    g = torch.Generator().manual_seed(N)
    lab = torch.stack([O.canonical_relabel(torch.randint(0, max(1, K - b % 3), (N,), generator=g))[0] for b in range(B)])
    ref = O.attention_mask(lab)
    got = msvit.attention_mask(lab.to(DEV))
    assert got.dtype == torch.bool and got.shape == ref.shape
    assert torch.equal(got.cpu(), ref)
    C = int(lab.max()) + 1
    assert torch.equal(msvit.attention_mask(lab.to(DEV), max_n_clusters=C).cpu(), ref)   # no host read of the labels


def test_ncut_eig_partial_convergence_request():
    # fixed K < ncut_dim: only the K pairs the k-means step reads must meet the tolerance (modeling_spectral.py:90);
    # they match the exact decomposition, and the solver stops early instead of chasing the noise-bulk pairs
    N, D, Kp, k = 256, 64, 3, 8
    B = 2
    x, _ = planted_tokens(B, N, D, Kp)
    A = torch.stack([O.affinity(x[b], "rbf", 3.0, default_scale(D)) for b in range(B)])
    deg = A.sum(-1)
    V_all, lam_all, it_all = F.ncut_eig(A.to(DEV), deg.to(DEV), k)
    V, lam, it = F.ncut_eig(A.to(DEV), deg.to(DEV), k, n_converge=Kp)
    assert int(it.max()) <= int(it_all.max())
    for b in range(B):
        Vref, lref, _ = O.ncut_eig(A[b].double(), k + 4)
        np.testing.assert_allclose(lam[b, :Kp].cpu().numpy(), lref[:Kp].numpy(), rtol=RTOL, atol=1e-6)
        check_eigvecs(V[b], lref.tolist(), Vref, Kp)


@pytest.mark.parametrize("shape", [(2, 3, 196, 8), (1, 2, 50, 5), (2, 1, 33, 12), (1, 4, 256, 20), (1, 2, 784, 39),
                                   (1, 1, 300, 64), (1, 2, 197, 8), (1, 1, 1024, 16), (64, 12, 196, 8),
                                   (1, 1, 1100, 8), (1, 1, 301, 5)])   # the last two take the two-pass route
def test_cluster_attention_stats_match_reference_restatement(shape):
    # compress_tokens_with_cluster_indices (msvitencoder.py:182-190): transmitter sums and receiver means
    B, H, N, C = shape
    g = torch.Generator().manual_seed(N + C)
    attn = torch.softmax(torch.randn(B, H, N, N, generator=g), dim=-1)
    lab = torch.randint(0, C, (B, N), generator=g)
    lab[0, lab[0] == C - 1] = 0          # an empty cluster
    tr_ref, rc_ref = O.cluster_attention_stats(attn.double(), lab, C)
    tr, rc = msvit.cluster_attention_stats(attn.to(DEV), lab.to(DEV), C)
    torch.testing.assert_close(tr.cpu().double(), tr_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rc.cpu().double(), rc_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tr.sum(-1).cpu(), torch.ones(B, H, N), rtol=1e-5, atol=1e-5)   # rows of a softmax


def test_cluster_attention_stats_labels_outside_the_cluster_range():
    # a query whose label is >= n_clusters joins no receiver row but keeps its transmitter row
    B, H, N, C = 2, 2, 64, 4
    g = torch.Generator().manual_seed(5)
    attn = torch.softmax(torch.randn(B, H, N, N, generator=g), dim=-1)
    lab = torch.randint(0, C + 2, (B, N), generator=g)
    tr_ref, rc_ref = O.cluster_attention_stats(attn.double(), lab, C)
    tr, rc = msvit.cluster_attention_stats(attn.to(DEV), lab.to(DEV), C)
    torch.testing.assert_close(tr.cpu().double(), tr_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rc.cpu().double(), rc_ref, rtol=1e-5, atol=1e-6)


def test_cluster_tokens_cosine_distance_matches_oracle():
    # model/clustering/modeling_spectral.py:62-69: the cosine variant of the NCUT affinity, whole path, bf16 tokens
    B, N, D, K = 4, 196, 768, 8
    x, planted = planted_tokens(B, N, D, K)
    out = msvit.cluster_tokens(x.bfloat16().to(DEV), ncut_dim=K, n_clusters=K, mode="cosine", gamma=0.5)
    xq = O.round_to_bf16(x)
    child, _, eigvals, _ = O.cluster_tokens(xq.double(), None, ncut_dim=K, n_clusters=K, mode="cosine", gamma=0.5)
    assert torch.equal(out.labels.cpu(), child)
    np.testing.assert_allclose(out.eigvals[:, 0].cpu().numpy(), eigvals[:, 0].numpy(), rtol=RTOL)
    for b in range(B):
        assert torch.equal(canon(planted[b]), out.labels[b].cpu())
