"""Dataset-level k-means (BASELINE.json configs[4]; reference call site model/clustering/modeling_spectral.py:254-256).

CPU part: the sharded host loop (`msvit.global_kmeans.lloyd` + one all-reduce of the packed sums|counts per
iteration) run by two gloo ranks with the oracle's arithmetic as the local step reproduces the single-process
oracle.  GPU part (`-m gpu`): every kernel of the iteration against the oracle, called through the C ABI.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msvit.global_kmeans import broadcast_init, lloyd
from msvit.sharding import gather_shards, shard_bounds
from oracle import ncut_oracle as O


def planted_features(n, D, k, seed=1212, noise=0.5):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(k, D, generator=g)
    lab = torch.randint(0, k, (n,), generator=g)
    x = centres[lab] + noise * torch.randn(n, D, generator=g)
    return x, lab


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


N_ROWS, DIM, KC, ITERS = 301, 16, 7, 4


def _oracle_local_step(x, C):
    """Packed [k, D+1] sums | counts of the rows of x against the centres C (the oracle's arithmetic)."""
    k, D = C.shape
    cn = (C * C).sum(-1)
    labels = torch.argmax(x @ C.T - 0.5 * cn[None, :], dim=1)
    packed = torch.zeros(k, D + 1, dtype=x.dtype)
    packed[:, :D].index_add_(0, labels, x)
    packed[:, D] = torch.bincount(labels, minlength=k).to(x.dtype)
    return packed, labels


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x_all, _ = planted_features(N_ROWS, DIM, KC)
        x_all = x_all.double()
        first, count = shard_bounds(N_ROWS, rank, world)
        x = x_all[first:first + count].contiguous()
        state = {"C": broadcast_init(x, KC).double(), "labels": None}

        def local_step():
            packed, state["labels"] = _oracle_local_step(x, state["C"])
            return packed

        def finalize(packed):
            cnt = packed[:, DIM]
            nz = cnt > 0
            state["C"][nz] = packed[nz, :DIM] / cnt[nz, None]

        lloyd(local_step, finalize, ITERS)
        labels_all = gather_shards(state["labels"], N_ROWS)
        if rank == 0:
            ret["C"] = state["C"]
            ret["labels"] = labels_all
    finally:
        dist.destroy_process_group()


def test_two_rank_lloyd_equals_single_process_oracle():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    x_all, _ = planted_features(N_ROWS, DIM, KC)
    C, labels, _ = O.global_kmeans(x_all.double(), KC, ITERS)
    assert torch.equal(ret["labels"], labels)
    torch.testing.assert_close(ret["C"], C, rtol=1e-12, atol=1e-12)   # fp64: only the summation order differs


def test_lloyd_single_rank_skips_the_collective():
    calls = []
    lloyd(lambda: calls.append("s") or torch.zeros(2, 3), lambda p: calls.append("f"), 3)
    assert calls == ["s", "f"] * 3


# ----------------------------------------------------------------------------------------------- GPU
DEV = "cuda:0"


def _operand(x, dtype):
    return O.round_to_bf16(x) if dtype == torch.bfloat16 else O.round_to_tf32(x)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(5000, 64, 37), (700, 768, 1000), (257, 32, 3), (4099, 128, 300)])
def test_assign_matches_oracle(dtype, shape):
    import msvit
    from msvit import _lib, ops
    n, D, k = shape
    x, _ = planted_features(n, D, min(k, 50))
    C = torch.randn(k, D, generator=torch.Generator().manual_seed(5))
    C[: min(k, n)] = x[: min(k, n)] * 0.9   # some centres close to data
    xo, Co = _operand(x, dtype).double(), _operand(C, dtype).double()
    score = (Co * Co).sum(-1)[None, :] - 2.0 * xo @ Co.T
    ref = torch.argmin(score, dim=1)
    lib = _lib.load()
    xg, Cg = x.to(dtype).to(DEV).contiguous(), C.to(dtype).to(DEV).contiguous()
    labels = torch.empty(n, dtype=torch.int32, device=DEV)
    best = torch.empty(n, dtype=torch.float32, device=DEV)
    code = _lib.F32 if dtype == torch.float32 else _lib.BF16
    _lib.check(lib.msvit_gkm_assign(ops._ptr(xg), code, ops._ptr(Cg), ops._ptr(labels), ops._ptr(best), n, k, D,
                                    torch.cuda.current_stream().cuda_stream), "assign")
    got = labels.cpu().long()
    # a different centre is acceptable only on a numerical tie of the scores
    diff = (got != ref).nonzero().flatten()
    s_got = score[torch.arange(n), got]
    s_ref = score[torch.arange(n), ref]
    assert torch.all((s_got - s_ref)[diff].abs() <= 1e-3 * (1.0 + s_ref[diff].abs())), f"{len(diff)} wrong labels"
    assert len(diff) <= max(1, n // 500)
    torch.testing.assert_close(best.cpu().double(), s_got, rtol=1e-3, atol=1e-2)


@pytest.mark.gpu
def test_sort_is_a_stable_counting_sort():
    from msvit import _lib, ops
    lib = _lib.load()
    for n, k in [(10000, 1000), (5, 3), (4097, 1), (70000, 17)]:
        g = torch.Generator().manual_seed(n)
        lab = torch.randint(0, k, (n,), generator=g, dtype=torch.int32)
        if n > 100:
            lab[lab == 2] = 0   # an empty label
        labg = lab.to(DEV)
        perm = torch.empty(n, dtype=torch.int32, device=DEV)
        seg = torch.empty(k + 1, dtype=torch.int32, device=DEV)
        wsb = int(lib.msvit_gkm_workspace_bytes(n, k))
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
        _lib.check(lib.msvit_gkm_sort(ops._ptr(labg), n, k, ops._ptr(perm), ops._ptr(seg), ops._ptr(ws), wsb,
                                      torch.cuda.current_stream().cuda_stream), "sort")
        ref_perm = torch.sort(lab.long(), stable=True).indices
        assert torch.equal(perm.cpu().long(), ref_perm)
        cnt = torch.bincount(lab.long(), minlength=k)
        assert torch.equal(seg.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(cnt, 0)]))


def _oracle_lloyd_with_operand_rounding(x, k, iters, dtype):
    """O.global_kmeans with the one deliberate difference of the CUDA path: the assignment reads the centres rounded
    to the tensor-core operand type (the fp32 master copy is what gets updated)."""
    C = x[:k].clone()
    labels = None
    for _ in range(iters):
        Cq = _operand(C.float(), dtype).double()
        score = (Cq * Cq).sum(-1)[None, :] - 2.0 * x @ Cq.T
        labels = torch.argmin(score, dim=1)
        sums = torch.zeros_like(C).index_add_(0, labels, x)
        counts = torch.bincount(labels, minlength=k)
        nz = counts > 0
        C[nz] = sums[nz] / counts[nz].to(C.dtype)[:, None]
    return C, labels, counts


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_global_kmeans_matches_oracle(dtype):
    import msvit
    n, D, k, iters = 6000, 64, 24, 6
    x, _ = planted_features(n, D, k, noise=0.3)
    xq = x.to(dtype)
    res = msvit.global_kmeans(xq.to(DEV), k, iters)
    C, labels, counts = _oracle_lloyd_with_operand_rounding(_operand(x, dtype).double(), k, iters, dtype)
    got = res.labels.cpu()
    assert (got != labels).float().mean() < 1e-3      # only rows on a numerical tie between two centres may differ
    torch.testing.assert_close(res.centroids.cpu().double(), C, rtol=1e-3, atol=1e-3)
    assert int(res.counts.sum()) == n
    # the plain oracle (no operand rounding of the centres) reaches the same clustering up to boundary rows
    C0, labels0, _ = O.global_kmeans(xq.double(), k, iters)
    assert (got != labels0).float().mean() < 2e-2
    again = msvit.global_kmeans(xq.to(DEV), k, iters)
    assert torch.equal(again.centroids, res.centroids) and torch.equal(again.labels, res.labels)   # bit-reproducible


@pytest.mark.gpu
def test_update_step_is_exact_and_keeps_empty_centres():
    import msvit
    from msvit.global_kmeans import GlobalKMeansPlan
    n, D, k = 3000, 96, 11
    x, _ = planted_features(n, D, 5)
    for dtype in (torch.float32, torch.bfloat16):
        xq = x.to(dtype)
        plan = GlobalKMeansPlan(n, D, k, dtype, DEV)
        init = torch.randn(k, D, generator=torch.Generator().manual_seed(3))
        init[7] = 1e3   # far from everything: stays empty and must keep its centre
        plan.set_centroids(init.to(DEV))
        packed = plan.local_step(xq.to(DEV)).clone()
        labels = plan.labels[:n].cpu().long()
        sums = torch.zeros(k, D, dtype=torch.float64).index_add_(0, labels, xq.double())
        cnt = torch.bincount(labels, minlength=k)
        assert cnt[7] == 0
        assert torch.equal(packed[:, D].cpu().long(), cnt)
        torch.testing.assert_close(packed[:, :D].cpu().double(), sums, rtol=1e-5, atol=1e-3)
        plan.finalize(packed)
        torch.testing.assert_close(plan.centroids[7].cpu(), init[7])
        nz = cnt > 0
        torch.testing.assert_close(plan.centroids[nz].cpu().double(), sums[nz] / cnt[nz, None], rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
def test_c5_full_size_sampled_parity():
    """BASELINE.json configs[4] at full size (1 000 000 x 768 bf16 rows, k = 1000), one Lloyd step: the assignment of a
    random sample of rows equals the fp32 argmin over the same bf16 operands, the member counts are exact, and the
    centroid sums of sampled centroids match a torch reduction of their member rows."""
    import msvit
    from msvit.global_kmeans import GlobalKMeansPlan
    dev = "cuda:0"
    n, D, k = 1_000_000, 768, 1000
    g = torch.Generator(device=dev).manual_seed(1212)
    centres = torch.randn(k, D, generator=g, device=dev)
    x = torch.empty(n, D, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, n, 1 << 16):
        r1 = min(n, r0 + (1 << 16))
        lab = torch.randint(0, k, (r1 - r0,), generator=g, device=dev)
        x[r0:r1] = (centres[lab] + 0.5 * torch.randn(r1 - r0, D, generator=g, device=dev)).bfloat16()
    plan = GlobalKMeansPlan(n, D, k, torch.bfloat16, dev)
    plan.set_centroids(x[:k].float())
    packed = plan.local_step(x).clone()
    torch.cuda.synchronize()
    labels = plan.labels[:n].long()
    cop = plan.centroids_op.float()                        # the bf16 operand the tensor cores read
    idx = torch.randint(0, n, (8192,), generator=g, device=dev)
    xs = x[idx].float()
    score = (cop * cop).sum(1)[None, :] - 2.0 * xs @ cop.T
    best = score.min(dim=1)
    got = score.gather(1, labels[idx][:, None])[:, 0]
    # the chosen centre attains the minimal score (ties / last-ulp differences allowed, a different centre is not)
    assert torch.all(got <= best.values + 1e-3 * best.values.abs().clamp_min(1.0))
    assert float((labels[idx] == best.indices).float().mean()) > 0.999
    counts = torch.bincount(labels, minlength=k)
    assert torch.equal(packed[:, D].round().long(), counts) and int(counts.sum()) == n
    for c in torch.randint(0, k, (12,), generator=g, device=dev).tolist():
        members = x[labels == c].float()
        torch.testing.assert_close(packed[c, :D], members.sum(0), rtol=2e-4, atol=2e-2)
    # perm is a stable sort by label: row ids ascending inside every label
    perm, seg = plan.perm[:n].long(), plan.seg_off.long()
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device=dev))
    c = int(torch.argmax(counts))
    run = perm[seg[c]:seg[c + 1]]
    assert torch.all(labels[run] == c) and torch.all(run[1:] > run[:-1])
