import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/multi-state-vit_b200")
import torch
from msvit.functional import ClusterPlan
from msvit.synthetic import default_scale, planted_tokens
B, N, D, K = 1024, 196, 768, 8
x, _ = planted_tokens(64, N, D, K); x = x.repeat(16, 1, 1).contiguous().cuda()
plan = ClusterPlan(B, N, D, torch.float32, "cuda", ncut_dim=K, n_clusters=K, scale=default_scale(D))
for _ in range(5): plan.run(x)
torch.cuda.synchronize()
def timed(with_events, steps=20):
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(ClusterPlan.STAGES) + 1)] for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record()
    for s in range(steps): plan.run(x, events=evs[s] if with_events else None)
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps
for _ in range(3):
    print(f"with events {timed(True):.4f} ms   without {timed(False):.4f} ms")
