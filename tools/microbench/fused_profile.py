"""Phase profile of ncut_fused_kernel (development aid).

    python tools/microbench/fused_profile.py build            (here, no GPU: builds _variants/libfused_prof.so with -DFUSED_PROFILE)
    python tools/microbench/fused_profile.py run [B] [bf16]   (on the GPU box)
"""
import ctypes, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multi-state-vit_b200", "csrc")
OUT = os.path.join(HERE, "_variants")
SO = os.path.join(OUT, "libfused_prof.so")

def build(extra):
    os.makedirs(OUT, exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DFUSED_PROFILE", "-shared", "-o", SO, os.path.join(CSRC, "ncut_fused.cu")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stderr[-2000:])
    assert r.returncode == 0

def run(argv):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
    import torch
    from msvit.synthetic import default_scale, planted_tokens
    B = int(argv[0]) if argv else 1024
    bf16 = len(argv) > 1 and argv[1] == "bf16"
    N, D, K = 196, 768, 8
    x, _ = planted_tokens(min(B, 64), N, D, K)
    x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1)[:B].contiguous().cuda()
    if bf16:
        x = x.bfloat16()
    lib = ctypes.CDLL(SO)
    fn = lib.msvit_ncut_fused
    fn.restype = ctypes.c_int
    P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    fn.argtypes = [P, I, P, P, P, P, P, L, I, I, I, I, F, F, I, I, F, F, I, P]
    rows = B * N
    deg = torch.zeros(rows, device="cuda"); U = torch.zeros(rows, 16, device="cuda"); H = torch.zeros(B, 256, device="cuda")
    iters = torch.zeros(B, dtype=torch.int32, device="cuda"); info = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def call():
        rc = fn(x.data_ptr(), 1 if bf16 else 0, deg.data_ptr(), U.data_ptr(), H.data_ptr(), iters.data_ptr(), info.data_ptr(),
                rows, B, N, D, 0, 3.0, default_scale(D), 16, 60, 2e-5, 0.0, 8, st)
        assert rc == 0, rc
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"fused kernel median {ts[len(ts)//2]:.4f} ms min {ts[0]:.4f}; iters mean {iters.float().mean():.2f} max {int(iters.max())} converged {int(info.sum())}/{B}")
    # load balance of the static segment -> CTA map against in-order dynamic dispatch (cost model: 42k + 8k * iterations)
    import heapq
    it = iters.cpu().tolist(); G = min(B, 148)
    cost = [42.0 + 8.0 * v for v in it]
    static = [sum(cost[c::G]) for c in range(G)]
    heap = [0.0] * G
    for c in cost:
        heapq.heappush(heap, heapq.heappop(heap) + c)
    print(f"k-cycles per CTA: static mean {sum(static) / G:.0f} max {max(static):.0f}; dynamic dispatch makespan {max(heap):.0f}")
    names = ["gram", "epilogue", "init", "uop", "product", "drain", "grams", "trigger", "chol", "subst", "output", "x1_cholesky_warp", "x2_ybar_wait"]
    buf = (ctypes.c_ulonglong * len(names))()
    lib.msvit_fused_profile(buf, 1)
    call(); torch.cuda.synchronize()
    lib.msvit_fused_profile(buf, 1)
    tot = sum(buf)
    print("cycles/segment: " + "  ".join(f"{n} {buf[i] / B:.0f}" for i, n in enumerate(names)) + f"  total {tot / B:.0f}")

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
