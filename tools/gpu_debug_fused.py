"""Development check of the fused NCut kernel against the two-kernel path and the oracle (run on the GPU box).

    python tools/gpu_debug_fused.py parity [B]      labels / eigenvalues / eigenvectors / degree, fused vs unfused vs oracle
    python tools/gpu_debug_fused.py time [B]        CUDA-event timing of both paths, per stage
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))

import torch  # noqa: E402

import msvit  # noqa: E402
from msvit.functional import ClusterPlan  # noqa: E402
from msvit.synthetic import default_scale, planted_tokens  # noqa: E402


def parity(B=8, dtype=torch.float32):
    from oracle import ncut_oracle as O
    N, D, K = 196, 768, 8
    x, _ = planted_tokens(B, N, D, K)
    xd = x.to("cuda:0").to(dtype)
    kw = dict(ncut_dim=K, n_clusters=K, scale=default_scale(D))
    ref = ClusterPlan(B, N, D, dtype, "cuda:0", fused=False, **kw)
    fus = ClusterPlan(B, N, D, dtype, "cuda:0", fused=True, **kw)
    o1 = ref.run(xd)
    torch.cuda.synchronize()
    print("unfused ok: iters", o1.iters.flatten().tolist()[:8])
    o2 = fus.run(xd)
    torch.cuda.synchronize()
    print("fused ok:   iters", o2.iters.flatten().tolist()[:8], "info", fus.info.tolist()[:8])
    print("deg    max rel diff", float(((o1.degree - o2.degree).abs() / o1.degree.abs()).max()))
    print("lam    unfused", o1.eigvals[0, 0].tolist())
    print("lam    fused  ", o2.eigvals[0, 0].tolist())
    print("lam    max rel diff", float(((o1.eigvals - o2.eigvals).abs() / o1.eigvals.abs()).max()))
    print("V      max abs diff", float((o1.eigvecs - o2.eigvecs).abs().max()))
    print("labels equal", bool(torch.equal(o1.labels, o2.labels)), " n_child", o2.n_child.flatten().tolist()[:8])
    print("pooled max abs diff", float((o1.pooled - o2.pooled).abs().max()), "counts equal", bool(torch.equal(o1.counts, o2.counts)))
    xin = O.round_to_bf16(x) if dtype == torch.bfloat16 else O.round_to_tf32(x)
    nb = min(B, 4)
    child, _, eigvals, _ = O.cluster_tokens(xin[:nb].double(), None, ncut_dim=K, n_clusters=K, scale=default_scale(D))
    print("oracle: labels equal (fused)", bool(torch.equal(o2.labels[:nb].cpu(), child)),
          " lam rel err", float(((o2.eigvals[:nb, 0].cpu().double() - eigvals[:, 0]).abs() / eigvals[:, 0].abs()).max()))
    for b in range(nb):
        A = O.affinity(xin[b].double(), "rbf", 3.0, default_scale(D))
        Vr, lr, dr = O.ncut_eig(A, K)
        ev = (o2.eigvecs[b].cpu().double() - Vr).norm(dim=0).max()
        ed = ((o2.degree[b].cpu().double() - dr).abs() / dr).max()
        print(f"  image {b}: eigvec col err {float(ev):.2e}  degree rel err {float(ed):.2e}")


def timing(B=1024, dtype=torch.float32):
    N, D, K = 196, 768, 8
    x, _ = planted_tokens(min(B, 64), N, D, K)
    x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1)[:B].contiguous()
    xd = x.to("cuda:0").to(dtype)
    kw = dict(ncut_dim=K, n_clusters=K, scale=default_scale(D))
    for name, fused in (("unfused", False), ("fused", True)):
        plan = ClusterPlan(B, N, D, dtype, "cuda:0", fused=fused, **kw)
        for _ in range(3):
            plan.run(xd)
        torch.cuda.synchronize()
        acc = [0.0] * 6
        tot = 0.0
        reps = 10
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            plan.run(xd, events=ev)
            torch.cuda.synchronize()
            for i in range(6):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
            tot += ev[0].elapsed_time(ev[6])
        print(f"{name:8s} {dtype}: step {tot / reps:.4f} ms  " +
              "  ".join(f"{s} {a / reps:.4f}" for s, a in zip(ClusterPlan.STAGES, acc)) +
              f"  iters mean {plan.iters.float().mean():.2f} max {int(plan.iters.max())}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else None
    dt = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
    if what == "parity":
        parity(B or 8, dt)
    else:
        timing(B or 1024, dt)
