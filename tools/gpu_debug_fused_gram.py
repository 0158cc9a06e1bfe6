"""Check G2 = Y^T D Y and H = U^T D Y of the fused kernel's first iteration against a CPU computation (development)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
from msvit.synthetic import default_scale, planted_tokens
from oracle import ncut_oracle as O
B, N, D, K = 2, 196, 768, 8
x, _ = planted_tokens(B, N, D, K)
def run(debug, max_iter):
    code = f'''
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "multi-state-vit_b200"))
os.environ["MSVIT_FUSED_DEBUG"] = "{debug}"
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
rh = run(128, 1)          # full-precision products, stop at the first iteration: H
rg = run(128 + 256, 1)    # ... G2
u0 = rh["U"][:N].double()
y = (A @ u0) / d[:, None]
Gref = y.T @ (d[:, None] * y)
Href = u0.T @ (d[:, None] * y)
G = rg["H"][0].view(16, 16).double(); H = rh["H"][0].view(16, 16).double()
torch.set_printoptions(precision=4, linewidth=200)
print("G rel err per row:", ((G - Gref).abs().max(1).values / Gref.abs().max(1).values).tolist())
print("H rel err per row:", ((H - Href).abs().max(1).values / Href.abs().max(1).values).tolist())
print("G[0,:6]", G[0, :6].tolist(), "ref", Gref[0, :6].tolist())
print("G[5,:6]", G[5, :6].tolist(), "ref", Gref[5, :6].tolist())
# ---- after one update: is U D-orthonormal?
ru = run(128 + 6, 2)
u = ru["U"][:N].double()
G1 = u.T @ (d[:, None] * u)
print("U^T D U diag:", [round(v, 4) for v in G1.diag().tolist()])
print("U^T D U row 0:", [round(v, 4) for v in G1[0].tolist()])
print("U^T D U row 1:", [round(v, 4) for v in G1[1].tolist()])
# what it should be: CholQR of y (twice)
def cholqr(Y):
    Gm = Y.T @ (d[:, None] * Y)
    L = torch.linalg.cholesky(Gm)
    return torch.linalg.solve_triangular(L, Y.T, upper=False).T, L
u1, L1 = cholqr(y)
print("ref L1 diag:", [round(v, 4) for v in L1.diag().tolist()])
u1b, _ = cholqr(u1)
print("|u - ref| per column:", [f"{float((u[:, c] - u1b[:, c]).norm() / u1b[:, c].norm()):.1e}" for c in range(16)])
print("kernel u col norms (D):", [round(float((d * u[:, c] ** 2).sum().sqrt()), 4) for c in range(16)])
# ---- the factor itself (no re-orthonormalisation): LT[a][c] = L[c][a]
rl = run(128 + 6 + 512 + 1024, 2)
LTk = rl["H"][0].view(16, 16).double()
print("kernel L diag:", [round(float(LTk[a, a]), 4) for a in range(16)])
print("kernel L col 0 (LT row 0):", [round(float(v), 3) for v in LTk[0].tolist()])
print("ref    L col 0:", [round(float(v), 3) for v in L1[:, 0].tolist()])
print("kernel L col 2 (LT row 2):", [round(float(v), 3) for v in LTk[2].tolist()])
print("ref    L col 2:", [round(float(v), 3) for v in L1[:, 2].tolist()])
