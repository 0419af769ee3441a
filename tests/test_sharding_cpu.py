"""World sharding + shared-gradient all-reduce on CPU with the gloo backend, world_size 2."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffdart_b200.sharding import allreduce_shared_gradient, shard, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 65536, 65537):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    g = torch.arange(10 * 3, dtype=torch.float64).reshape(10, 3)  # the full per-world gradient
    mine = shard(g, rank, world_size)
    total = allreduce_shared_gradient(mine)
    if rank == 0:
        np.save(out, total.numpy())
    dist.destroy_process_group()


def test_allreduce_shared_gradient_gloo(tmp_path):
    out = str(tmp_path / "total.npy")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    want = torch.arange(10 * 3, dtype=torch.float64).reshape(10, 3).sum(0).numpy()
    assert np.array_equal(np.load(out), want)
