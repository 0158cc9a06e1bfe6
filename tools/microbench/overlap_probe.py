"""Does ritz_kmeans / pool co-run with ncut_fused_kernel when the latter leaves shared memory free?  (development aid)
    MSVIT_FUSED_RESERVE_KB=24 python tools/microbench/overlap_probe.py
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
from msvit import _lib, ops
from msvit.functional import ClusterPlan
from msvit.synthetic import default_scale, planted_tokens

B, N, D, K = 1024, 196, 768, 8
x, _ = planted_tokens(64, N, D, K)
x = x.repeat(B // 64, 1, 1).contiguous().cuda()
plan = ClusterPlan(B, N, D, torch.float32, "cuda", ncut_dim=K, n_clusters=K, scale=default_scale(D), fused=True)
out = plan.run(x); torch.cuda.synchronize()
lib = _lib.load()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
labels = out.labels.clone()
pooled = torch.empty(B, K, D, device="cuda"); counts = torch.empty(B, K, dtype=torch.int32, device="cuda")

def fused(st):
    _lib.check(lib.msvit_ncut_fused(ops._ptr(x), _lib.F32, ops._ptr(plan.deg), ops._ptr(plan.U), ops._ptr(plan.H), ops._ptr(plan.iters),
               ops._ptr(plan.info), B * N, B, N, D, plan.mode, plan.gamma, plan.scale, 16, plan.eig_iters, plan.eig_tol, plan.lam_floor, plan.n_converge, st), "fused")
def ritz(st):
    _lib.check(lib.msvit_ritz_kmeans(ops._ptr(plan.U2), ops._ptr(plan.H2), ops._ptr(plan.info2), ops._ptr(plan.deg2), ops._ptr(plan.V), ops._ptr(plan.lam),
               None, ops._ptr(plan.child), ops._ptr(plan.n_child), B * N, B, N, K, 16, K, K, 0.0, 100, 0, st), "ritz")
def pool(st):
    _lib.check(lib.msvit_pool(ops._ptr(x), _lib.F32, ops._ptr(labels), ops._ptr(pooled), ops._ptr(counts), B, N, D, K, st), "pool")

plan.U2, plan.H2, plan.info2, plan.deg2 = plan.U.clone(), plan.H.clone(), plan.info.clone(), plan.deg.clone()

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]

cur = torch.cuda.current_stream()
def serial():
    st = cur.cuda_stream; fused(st); ritz(st); pool(st)
def only_fused(): fused(cur.cuda_stream)
def only_post():
    st = cur.cuda_stream; ritz(st); pool(st)
def overlapped():
    ev0 = torch.cuda.Event(); ev0.record(cur)
    sA.wait_event(ev0); sB.wait_event(ev0)
    fused(sA.cuda_stream)
    ritz(sB.cuda_stream); pool(sB.cuda_stream)
    eA, eB = torch.cuda.Event(), torch.cuda.Event(); eA.record(sA); eB.record(sB)
    cur.wait_event(eA); cur.wait_event(eB)
print("reserve KB", os.environ.get("MSVIT_FUSED_RESERVE_KB"))
print(f"fused {timed(only_fused):.4f}  post {timed(only_post):.4f}  serial {timed(serial):.4f}  overlapped {timed(overlapped):.4f} ms")
