"""Times the affinity kernel on the C2 shape: fp32 / bf16 tokens, with and without the affinity output (dev aid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
from msvit import _lib, ops

B, N, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 196, 768)
dev = "cuda:0"
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for dtype in (torch.float32, torch.bfloat16):
    x = torch.randn(B * N, D, device=dev).to(dtype)
    lda = ops.lda_of(N)
    A = torch.empty(B * N * lda, device=dev)
    deg = torch.empty(B * N, device=dev)
    code = _lib.F32 if dtype == torch.float32 else _lib.BF16
    for want in (True, False):
        def call():
            rc = lib.msvit_affinity_degree(x.data_ptr(), code, A.data_ptr() if want else None, deg.data_ptr(), B * N, B, N, D, 0,
                                           3.0, D / 4.0, None, None, st)
            assert rc == 0
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        byts = x.numel() * x.element_size() + (A.numel() * 4 if want else 0)
        print(f"{dtype} A={'yes' if want else 'no '}: {ts[5]:.4f} ms  {byts / ts[5] / 1e6:.0f} GB/s  {2.0 * B * N * N * D / ts[5] / 1e9:.0f} TFLOP/s", flush=True)
