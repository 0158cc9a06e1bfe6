"""Phase profile of ritz_kmeans_kernel (development aid).
    python tools/microbench/km_profile.py build      (here)      python tools/microbench/km_profile.py run [B]   (GPU box)
"""
import ctypes, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multi-state-vit_b200", "csrc"); OUT = os.path.join(HERE, "_variants"); SO = os.path.join(OUT, "libkm_prof.so")

def build(extra):
    os.makedirs(OUT, exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DKM_PROFILE", "-shared", "-o", SO, os.path.join(CSRC, "kmeans.cu")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True); print(r.stderr[-1500:]); assert r.returncode == 0

def run(argv):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
    import torch
    from msvit.functional import ClusterPlan
    from msvit.synthetic import default_scale, planted_tokens
    B = int(argv[0]) if argv else 1024
    N, D, K = 196, 768, 8
    x, _ = planted_tokens(min(B, 64), N, D, K)
    x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1)[:B].contiguous().cuda()
    plan = ClusterPlan(B, N, D, torch.float32, "cuda", ncut_dim=K, n_clusters=K, scale=default_scale(D), fused=True)
    plan.run(x); torch.cuda.synchronize()
    lib = ctypes.CDLL(SO)
    fn = lib.msvit_ritz_kmeans; fn.restype = ctypes.c_int
    P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    fn.argtypes = [P, P, P, P, P, P, P, P, P, L, I, I, I, I, I, I, F, I, I, P]
    st = torch.cuda.current_stream().cuda_stream
    def call():
        rc = fn(plan.U.data_ptr(), plan.H.data_ptr(), plan.info.data_ptr(), plan.deg.data_ptr(), plan.V.data_ptr(), plan.lam.data_ptr(),
                None, plan.child.data_ptr(), plan.n_child.data_ptr(), B * N, B, N, K, 16, K, K, 0.0, 100, 0, st)
        assert rc == 0, rc
    for _ in range(3): call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); print(f"ritz_kmeans median {ts[5]:.4f} ms min {ts[0]:.4f}")
    names = ["jacobi", "rotate", "seed", "lloyd", "relabel"]
    buf = (ctypes.c_ulonglong * len(names))()
    lib.msvit_km_profile(buf, 1); call(); torch.cuda.synchronize(); lib.msvit_km_profile(buf, 1)
    print("cycles/segment: " + "  ".join(f"{n} {buf[i] / B:.0f}" for i, n in enumerate(names)) + f"  total {sum(buf) / B:.0f}")

if __name__ == "__main__":
    build(sys.argv[2:]) if sys.argv[1] == "build" else run(sys.argv[2:])
