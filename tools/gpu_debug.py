"""Stage-by-stage diagnostics on a GPU box (development aid, not part of the test suite).

    python tools/gpu_debug.py <stage>     stage in: pool kmeans eig affinity_bf16 affinity_f32 e2e time
Each stage prints error statistics against the oracle instead of asserting, so one run shows everything.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))

import torch  # noqa: E402

import msvit  # noqa: E402
from msvit import functional as F, ops  # noqa: E402
from msvit.synthetic import default_scale, planted_tokens  # noqa: E402
from oracle import ncut_oracle as O  # noqa: E402

DEV = "cuda:0"


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / (b.abs() + 1e-12)).max()), float((a - b).abs().max())


def stage_pool():
    B, N, D, K = 4, 196, 768, 8
    x = torch.randn(B, N, D)
    lab = torch.randint(0, K, (B, N))
    for dt in (torch.float32, torch.bfloat16):
        p, c = msvit.pool(x.to(dt).to(DEV), lab.to(DEV), K)
        rp, rc = O.pool(x.to(dt).double(), lab, K)
        print("pool", dt, "counts equal", torch.equal(c.cpu(), rc), "err", relerr(p, rp))


def stage_kmeans():
    B, N, D, K = 4, 196, 768, 8
    x, _ = planted_tokens(B, N, D, K)
    Vs, ds, ls = [], [], []
    for b in range(B):
        V, lam, deg = O.ncut_eig(O.affinity(x[b], "rbf", 3.0, default_scale(D)), K)
        l, _, C = O.kmeans(V, K, weight=deg)
        Vs.append(V); ds.append(deg); ls.append(l)
    labels, n_child, cen = F.kmeans(torch.stack(Vs).to(DEV), K, weight=torch.stack(ds).to(DEV))
    print("kmeans labels equal", torch.equal(labels.cpu(), torch.stack(ls)), "n_child", n_child.cpu().tolist())


def stage_eig():
    for (N, D, Kp, k) in [(196, 768, 8, 8), (576, 1024, 16, 16), (10, 16, 2, 8)]:
        B = 2
        x, _ = planted_tokens(B, N, D, Kp)
        lda = ops.lda_of(N)
        A = torch.zeros(B, N, lda)
        refs = []
        for b in range(B):
            Ab = O.affinity(x[b], "rbf", 3.0, default_scale(D))
            A[b, :, :N] = Ab
            refs.append(O.ncut_eig(Ab.double(), min(k, N)))
        deg = A.sum(-1)
        t = time.time()
        V, lam, iters = F.ncut_eig(A.to(DEV), deg.to(DEV), k)
        torch.cuda.synchronize()
        print(f"eig N={N} k={k}: iters {iters.cpu().tolist()} time {time.time() - t:.3f}s")
        for b in range(B):
            Vr, lr, _ = refs[b]
            kk = min(k, N)
            print("   lam", lam[b, :kk].cpu().numpy().round(5).tolist())
            print("   ref", lr[:kk].numpy().round(5).tolist())
            print("   subspace dist", O.subspace_distance(V[b, :, :min(kk, Kp)].double().cpu(), Vr[:, :min(kk, Kp)]),
                  "col errs", [round(float(torch.linalg.norm(V[b, :, j].double().cpu() - Vr[:, j])), 5) for j in range(min(kk, Kp))])


def stage_affinity(dtype):
    for (B, N, D) in [(2, 196, 768), (2, 64, 64), (1, 300, 128), (1, 576, 1024), (2, 37, 24)]:
        x, _ = planted_tokens(B, N, D, 4)
        xq = O.round_to_bf16(x) if dtype == torch.bfloat16 else O.round_to_tf32(x)
        try:
            A, deg = F.affinity(x.to(dtype).to(DEV), "rbf", 3.0, default_scale(D))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("affinity", dtype, (B, N, D), "FAILED", repr(e))
            raise
        for b in range(B):
            ref = O.affinity(xq[b].double(), "rbf", 3.0, default_scale(D))
            e = relerr(A[b, :, :N], ref)
            ed = relerr(deg[b], ref.sum(-1))
            print("affinity", dtype, (B, N, D), "img", b, "A err", e, "deg err", ed)
            if e[0] > 1e-3:
                print("   got", A[b, :3, :6].cpu().numpy())
                print("   ref", ref[:3, :6].numpy())
                print("   got tail", A[b, -2:, -6:].cpu().numpy())
                print("   ref tail", ref[-2:, -6:].numpy())
                bad = ((A[b, :, :N].cpu().double() - ref).abs() / ref > 1e-3)
                print("   bad fraction", float(bad.float().mean()), "bad rows", bad.any(1).nonzero().flatten()[:20].tolist(),
                      "bad cols", bad.any(0).nonzero().flatten()[:20].tolist())


def stage_e2e():
    B, N, D, K = 8, 196, 768, 8
    x, planted = planted_tokens(B, N, D, K)
    for dt in (torch.float32, torch.bfloat16):
        out = msvit.cluster_tokens(x.to(dt).to(DEV), ncut_dim=K, n_clusters=K, scale=default_scale(D))
        torch.cuda.synchronize()
        xq = O.round_to_bf16(x) if dt == torch.bfloat16 else O.round_to_tf32(x)
        child, _, lam, _ = O.cluster_tokens(xq.double(), None, ncut_dim=K, n_clusters=K, scale=default_scale(D))
        print("e2e", dt, "labels equal", torch.equal(out.labels.cpu(), child), "iters", out.iters.flatten().cpu().tolist())
        print("   lam err", relerr(out.eigvals[:, 0], lam[:, 0]))
        rp, rc = O.pool(x.to(dt).double(), child, K)
        print("   pooled err", relerr(out.pooled, rp), "counts equal", torch.equal(out.counts.cpu(), rc))


def stage_time():
    B, N, D, K = 1024, 196, 768, 8
    x, _ = planted_tokens(B, N, D, K)
    for dt in (torch.float32, torch.bfloat16):
        xg = x.to(dt).to(DEV)
        flat = xg.view(B * N, D)
        s = default_scale(D)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for rep in range(3):
            ev[0].record()
            A, deg = ops.affinity_degree(flat, B, N, 0, 3.0, s, None, None, B * N * N, True)
            ev[1].record()
            V, lam, iters = ops.ncut_eig(A, deg, B, N, K, 16, 60, 2e-5, None, None)
            ev[2].record()
            lab, nch, _ = ops.kmeans(V, lam, deg, None, B, N, K, 0.0, 100, None)
            ev[3].record()
            child = ops.compose_labels(lab, nch, None, None, B, N, 1)
            ev[4].record()
            pooled, counts = ops.pool(xg, child, K)
            ev[5].record()
            torch.cuda.synchronize()
        names = ["affinity", "eig", "kmeans", "compose", "pool"]
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
        print("time", dt, {n: round(m, 3) for n, m in zip(names, ms)}, "total ms", round(sum(ms), 3),
              "img/s", round(B / sum(ms) * 1e3), "iters max", int(iters.max()), "mean", float(iters.float().mean()))


if __name__ == "__main__":
    st = sys.argv[1]
    {"pool": stage_pool, "kmeans": stage_kmeans, "eig": stage_eig,
     "affinity_bf16": lambda: stage_affinity(torch.bfloat16), "affinity_f32": lambda: stage_affinity(torch.float32),
     "e2e": stage_e2e, "time": stage_time}[st]()
