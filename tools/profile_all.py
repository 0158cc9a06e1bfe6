"""One launch of every kernel of the library at its benchmark shape, bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r2_all python tools/profile_all.py
(run it once without ncu first).  Shapes: C2 (fused pair + pool), C3 (affinity, ncut_eig, kmeans, pool), one hierarchical
level with 4 parents per image (segments, gather, compose), C5 (one Lloyd iteration of the global k-means), the
attention mask and the one-pass attention statistics, the axis-aligned discretisation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
import msvit
from msvit.functional import ClusterPlan
from msvit.global_kmeans import GlobalKMeansPlan, broadcast_init
from msvit.synthetic import default_scale, planted_tokens

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)

def tokens(B, N, D, K):
    x, _ = planted_tokens(min(B, 64), N, D, K)
    return x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1)[:B].contiguous().to(dev)

x2 = tokens(1024, 196, 768, 8)
p2 = ClusterPlan(1024, 196, 768, torch.float32, dev, ncut_dim=8, n_clusters=8, scale=default_scale(768), fused=True)
p2k = ClusterPlan(1024, 196, 768, torch.float32, dev, ncut_dim=8, n_clusters=8, scale=default_scale(768), fused=True,
                  discretise="axis_align")
x3 = tokens(512, 576, 1024, 16)
p3 = ClusterPlan(512, 576, 1024, torch.float32, dev, ncut_dim=16, n_clusters=16, scale=default_scale(1024))
x4 = tokens(64, 1024, 768, 4)
p4a = ClusterPlan(64, 1024, 768, torch.float32, dev, ncut_dim=8, n_clusters=4, scale=default_scale(768))
p4b = ClusterPlan(64, 1024, 768, torch.float32, dev, ncut_dim=8, n_clusters=4, scale=default_scale(768), n_parents=4)
n5, D5, k5 = 1_000_000, 768, 1000
g = torch.Generator(device=dev).manual_seed(1212)
cent = torch.randn(k5, D5, device=dev, generator=g)
x5 = (cent[torch.randint(0, k5, (n5,), device=dev, generator=g)] + 0.5 * torch.randn(n5, D5, device=dev, generator=g)).bfloat16()
p5 = GlobalKMeansPlan(n5, D5, k5, torch.bfloat16, dev)
p5.set_centroids(broadcast_init(x5, k5))
ci = torch.randint(0, 8, (256, 196), device=dev)
attn = torch.softmax(torch.randn(64, 12, 196, 196, device=dev), dim=-1)

def everything():
    o2 = p2.run(x2)
    p2k.run(x2)
    p3.run(x3)
    parents = p4a.run(x4).labels
    p4b.run(x4, parents)
    p5.finalize(p5.local_step(x5))
    msvit.attention_mask(ci, 8)
    msvit.cluster_attention_stats(attn, ci[:64], 8)
    return o2

everything(); everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
