"""Build and time compile-time variants of the eigensolver kernel (development aid).

    python tools/microbench/eig_variants.py build  "T=4,MINB=4" "T=3,MINB=4,RR_EVERY=6" ...   (here, no GPU)
    python tools/microbench/eig_variants.py run [B N D K k]                                    (on the GPU box)

`build` compiles ncut_eig.cu once per variant into tools/microbench/_variants/libeig_<i>.so (macro EIG_<NAME>=<v>);
`run` computes the C2 affinity once with the shipped library, then for every variant times msvit_ncut_eig with
CUDA events (L2 flushed by the 157 MB affinity working set itself) and compares eigenvalues / eigenvectors with
the shipped build.
"""
import ctypes
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multi-state-vit_b200", "csrc")
OUT = os.path.join(HERE, "_variants")


def build(specs):
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.remove(os.path.join(OUT, f))
    meta = []
    for i, spec in enumerate(specs):
        defs = [f"-DEIG_{kv.split('=')[0]}={kv.split('=')[1]}" if "=" in kv else f"-DEIG_{kv}" for kv in spec.split(",") if kv]
        so = os.path.join(OUT, f"libeig_{i}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
               "--expt-relaxed-constexpr", "-Xptxas", "-v", "-shared", "-o", so, os.path.join(CSRC, "ncut_eig.cu")] + defs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr)
            raise SystemExit(1)
        regs = [l for l in r.stderr.splitlines() if "Used" in l or "spill" in l]
        k128 = [j for j, l in enumerate(r.stderr.splitlines()) if "ILi1ELi128E" in l and "Compiling" in l]
        info = ""
        lines = r.stderr.splitlines()
        for j in k128:
            info = " | ".join(x.strip() for x in lines[j + 1:j + 4])
        print(f"[{i}] {spec}: {info}")
        meta.append(spec)
    json.dump(meta, open(os.path.join(OUT, "variants.json"), "w"))


def run(argv):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
    import torch
    from msvit import _lib, functional as F, ops
    from msvit.synthetic import default_scale, planted_tokens
    B, N, D, K, k = (int(a) for a in argv) if argv else (1024, 196, 768, 8, 8)
    dev = "cuda:0"
    x, _ = planted_tokens(B, N, D, K)
    A, deg = F.affinity(x.to(dev), "rbf", 3.0, default_scale(D))
    A = A.contiguous().view(-1)
    deg = deg.contiguous().view(-1)
    block = F.default_block(k)
    ref = None
    meta = json.load(open(os.path.join(OUT, "variants.json")))
    libs = [("shipped", os.path.join(CSRC, "libmsvit.so"))] + [(s, os.path.join(OUT, f"libeig_{i}.so")) for i, s in enumerate(meta)]
    st = torch.cuda.current_stream().cuda_stream
    for name, path in libs:
        lib = ctypes.CDLL(path)
        fn = lib.msvit_ncut_eig
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64] + [ctypes.c_int] * 5 + [ctypes.c_float] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 3
        V = torch.empty(B * N, k, device=dev)
        lam = torch.empty(B, k, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)

        def call():
            rc = fn(A.data_ptr(), deg.data_ptr(), V.data_ptr(), lam.data_ptr(), iters.data_ptr(), B * N, B, N, k, block, 60,
                    float(os.environ.get("EIG_TOL", "2e-5")), 0.0, int(os.environ.get("EIG_NCONV", "0")), None, None, st)
            assert rc == 0, rc
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        msg = f"{name:40s} median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}  iters mean {iters.float().mean():.2f} max {int(iters.max())}"
        if ref is None:
            ref = (V.clone(), lam.clone())
        else:
            msg += f"  dlam {float((lam - ref[1]).abs().max()):.2e}  dV {float((V - ref[0]).abs().max()):.2e}"
        print(msg, flush=True)
        if hasattr(lib, "msvit_eig_profile"):
            buf = (ctypes.c_ulonglong * 11)()
            lib.msvit_eig_profile(buf, 1)
            call()
            torch.cuda.synchronize()
            lib.msvit_eig_profile(buf, 1)
            names = ["init", "matvec", "grams", "trigger", "jacobi", "rotate", "chol", "orth", "output", "(fact", "inv)"]
            tot = sum(buf[:9])
            print("    cycles/segment: " + "  ".join(f"{n} {buf[i] / B:.0f}" for i, n in enumerate(names)) + f"  total {tot / B:.0f}",
                  flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
