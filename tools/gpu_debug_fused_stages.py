"""Bisect the fused kernel: run msvit_ncut_fused alone at a debug stage (env MSVIT_FUSED_DEBUG) and synchronise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
from msvit import _lib, ops
from msvit.synthetic import default_scale, planted_tokens
B, N, D, K = 2, 196, 768, 8
x, _ = planted_tokens(B, N, D, K)
xd = x.cuda()
lib = _lib.load()
rows = B * N
deg = torch.zeros(rows, device="cuda"); U = torch.zeros(rows, 16, device="cuda"); H = torch.zeros(B, 256, device="cuda")
iters = torch.zeros(B, dtype=torch.int32, device="cuda"); info = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rc = lib.msvit_ncut_fused(xd.data_ptr(), 0, deg.data_ptr(), U.data_ptr(), H.data_ptr(), iters.data_ptr(), info.data_ptr(),
                          rows, B, N, D, 0, 3.0, default_scale(D), 16, 60, 2e-5, 0.0, 8, st)
print("rc", rc)
torch.cuda.synchronize()
print("stage", os.environ.get("MSVIT_FUSED_DEBUG"), "ok; deg[:4]", deg[:4].tolist(), "iters", iters.tolist(), "U[0,:4]", U[0, :4].tolist())
if os.environ.get("MSVIT_FUSED_DEBUG") == "1":
    from oracle import ncut_oracle as O
    A = O.affinity(O.round_to_tf32(x[0]).double(), "rbf", 3.0, default_scale(D))
    print("   oracle deg[:4]", A.sum(-1)[:4].tolist())
