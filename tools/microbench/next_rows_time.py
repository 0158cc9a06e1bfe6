"""Times the SURVEY 8f kernels built this round (attention mask, cluster-compressed attention statistics)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
import msvit

dev = "cuda:0"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B, N, C = 1024, 196, 8
lab = torch.randint(0, C, (B, N), device=dev)
ms = timeit(lambda: msvit.attention_mask(lab, max_n_clusters=C))
L = 2 * C + N
print(f"attention_mask B={B} N={N} C={C}: {ms:.4f} ms, {B * L * L / ms / 1e6:.0f} GB/s written")
for Bh, H, N, C in [(128, 12, 196, 8), (64, 12, 256, 16), (16, 12, 784, 39), (8, 12, 1100, 8)]:
    attn = torch.softmax(torch.randn(Bh, H, N, N, device=dev), -1)
    labh = torch.randint(0, C, (Bh, N), device=dev)
    ms = timeit(lambda: msvit.cluster_attention_stats(attn, labh, C))
    byts = attn.numel() * 4
    route = "one pass" if N <= 256 and C <= 16 else "two passes"
    print(f"cluster_attention_stats B={Bh} H={H} N={N} C={C}: {ms:.4f} ms, {byts / 1e6:.0f} MB of attention, {route} "
          f"({byts / ms / 1e6:.0f} GB/s of attention per unit time)")
    del attn
