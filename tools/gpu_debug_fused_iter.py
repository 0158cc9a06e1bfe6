"""Check the fused iteration's invariant y = D^-1 A u after a given number of updates (development)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
from msvit import _lib
from msvit.synthetic import default_scale, planted_tokens
from oracle import ncut_oracle as O
B, N, D, K = 2, 196, 768, 8
x, _ = planted_tokens(B, N, D, K)
def run(stage, max_iter):
    code = f'''
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "multi-state-vit_b200"))
os.environ["MSVIT_FUSED_DEBUG"] = "{stage}"
import torch
from msvit import _lib
from msvit.synthetic import default_scale, planted_tokens
B, N, D, K = {B}, {N}, {D}, {K}
x, _ = planted_tokens(B, N, D, K)
xd = x.cuda(); lib = _lib.load(); rows = B * N
deg = torch.zeros(rows, device="cuda"); U = torch.zeros(rows, 16, device="cuda"); H = torch.zeros(B, 256, device="cuda")
iters = torch.zeros(B, dtype=torch.int32, device="cuda"); info = torch.zeros(B, dtype=torch.int32, device="cuda")
rc = lib.msvit_ncut_fused(xd.data_ptr(), 0, deg.data_ptr(), U.data_ptr(), H.data_ptr(), iters.data_ptr(), info.data_ptr(), rows, B, N, D, 0, 3.0, default_scale(D), 16, {max_iter}, 2e-5, 0.0, 8, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
torch.save(dict(U=U.cpu(), deg=deg.cpu(), H=H.cpu(), iters=iters.cpu()), "/tmp/fused_dbg.pt")
'''
    subprocess.run([sys.executable, "-c", code], check=True)
    return torch.load("/tmp/fused_dbg.pt")
A = O.affinity(O.round_to_tf32(x[0]).double(), "rbf", 3.0, default_scale(D))
d = A.sum(-1)
for upd in (1, 2, 3, 4, 5):
    ru = run(6, upd + 1)
    ry = run(7, upd + 1)
    u = ru["U"][:N].double(); y = ry["U"][:N].double()
    G = u.T @ (d[:, None] * u)
    yref = (A @ u) / d[:, None]
    print(f"after {upd} updates: |U^T D U - I| = {float((G - torch.eye(16)).abs().max()):.2e}   |y - D^-1 A u| / |y| per column:",
          " ".join(f"{float((y[:, c] - yref[:, c]).norm() / yref[:, c].norm()):.1e}" for c in range(16)), " iters", ru["iters"].tolist())
