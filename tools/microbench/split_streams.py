"""Experiment: one C2 step as 1 / 2 / 4 sub-batches on separate streams (does kernel-level overlap hide the eigensolver's
partial second wave and the latency-bound small kernels?)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
from msvit.functional import ClusterPlan
from msvit.synthetic import default_scale, planted_tokens

B, N, D, K = 1024, 196, 768, 8
dev = torch.device("cuda", 0)
x = planted_tokens(B, N, D, K)[0].to(dev)
for parts in (1, 2, 3, 4, 8):
    sizes = [B // parts + (1 if i < B % parts else 0) for i in range(parts)]
    plans = [ClusterPlan(s, N, D, torch.float32, dev, ncut_dim=K, n_clusters=K, scale=default_scale(D)) for s in sizes]
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    offs = [sum(sizes[:i]) for i in range(parts)]
    xs = [x[o:o + s].contiguous() for o, s in zip(offs, sizes)]
    cur = torch.cuda.current_stream(dev)

    def step():
        ev = torch.cuda.Event()
        ev.record(cur)
        for p, st, xi in zip(plans, streams, xs):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                p.run(xi)
        for st in streams:
            cur.wait_stream(st)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"parts={parts}: {ms:.4f} ms/step  {B / ms * 1e3:,.0f} images/s", flush=True)
