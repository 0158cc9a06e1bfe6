"""World-size-2 gloo tests (CPU) of the batch-sharding host logic used by bench.py under torchrun:
contiguous partition, rank-local input generation, result gather, max-over-ranks timing.  The compute inside
each shard is the CPU oracle here (no GPU in this test); the GPU parity tests check that a shard clustered
alone gives bit-identical labels to the same images inside the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msvit.sharding import gather_shards, max_over_ranks, shard_bounds
from msvit.synthetic import default_scale, planted_tokens
from oracle import ncut_oracle as O

B, N, D, K = 5, 48, 32, 3  # odd batch: shards of 3 and 2 images


def test_shard_bounds_partition():
    for total in (0, 1, 5, 8, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard_bounds(B, rank, world)
        x, _ = planted_tokens(count, N, D, K, first=first)  # every rank generates exactly its own images
        child, _, lam, _ = O.cluster_tokens(x.double(), None, ncut_dim=4, n_clusters=K, scale=default_scale(D))
        pooled, counts = O.pool(x.double(), child, K)
        labels_all = gather_shards(child, B)
        pooled_all = gather_shards(pooled, B)
        slow = max_over_ranks(10.0 + rank, torch.device("cpu"))
        if rank == 0:
            ret["labels"] = labels_all
            ret["pooled"] = pooled_all
            ret["slow"] = slow
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    x, _ = planted_tokens(B, N, D, K)
    child, _, _, _ = O.cluster_tokens(x.double(), None, ncut_dim=4, n_clusters=K, scale=default_scale(D))
    pooled, _ = O.pool(x.double(), child, K)
    assert torch.equal(ret["labels"], child)          # bit-identical labels, 1 rank vs 2 ranks
    assert torch.equal(ret["pooled"], pooled)
    assert ret["slow"] == 11.0                        # step time = slowest rank
