"""Pinned host -> device bandwidth per GPU with all ranks copying at once (the ceiling of every end-to-end number).

    python tools/h2d_probe.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe.py
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))
import torch
import torch.distributed as dist
from msvit.sharding import bind_host_thread_to_gpu, max_over_ranks

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = {}
for bind in (False, True):
    cores = bind_host_thread_to_gpu(local) if bind else None
    nbytes = 617 * 1000 * 1000
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / 10
    out["bound_to_gpu_numa" if bind else "default_affinity"] = {"GB/s_per_gpu": round(nbytes / ms / 1e6, 1), "cores": None if cores is None else len(cores)}
    del h, d
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_617MB_pinned": out}))
if world > 1:
    dist.destroy_process_group()
