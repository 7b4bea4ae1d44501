"""CPU, world_size 2 over gloo: the data-parallel plumbing (batch sharding + ONE all-reduce over
the flat gradient buffer + identical Adam update on every replica)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pn2_b200.train_step import allreduce_flat, shard_batch
    # each rank owns a contiguous shard of the global batch
    gb = np.arange(8 * 5, dtype=np.float32).reshape(8, 5)
    mine = shard_batch(gb, rank, world)
    assert mine.shape[0] == 4 and mine[0, 0] == rank * 4 * 5
    # "gradient" = a function of the shard; the all-reduced mean must equal the global mean
    g = torch.tensor(mine.sum(0))
    flat = torch.cat([g, torch.full((3,), float(rank + 1))])
    scale = allreduce_flat(flat, world)
    res = flat * scale
    exp = np.concatenate([gb.sum(0) / world, np.full(3, (1 + 2) / 2.0)])
    np.testing.assert_allclose(res.numpy(), exp, rtol=1e-6)
    out[rank] = 1
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}
