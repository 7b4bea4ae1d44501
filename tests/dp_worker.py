"""Worker of tests/test_train_step_gpu.py::test_data_parallel_two_ranks_nccl (one process per GPU, NCCL).

Checks, on every rank: (1) after the first step's broadcast all ranks hold rank 0's weights although each
rank initialised from a different seed; (2) the all-reduced flat gradient times 1/world equals the mean of
the per-rank shard gradients (gathered before the reduction); (3) after 2 steps (eager, then CUDA-graph
replay) the flat weights are bit-identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import pn2_b200  # noqa: F401
    from pn2_b200 import train_step
    from pn2_b200.train_step import Trainer, shard_batch
    from test_train_step_gpu import HP_SMALL, small_batches
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    (pc, labels, smpw), = small_batches(1, b=2 * world)
    mine = [torch.as_tensor(shard_batch(x, rank, world)).to(dev) for x in (pc, labels, smpw)]
    tr = Trainer(dict(HP_SMALL, batch_size=2 * world), 9, device=dev, seed=rank, world_size=world)

    # (2) intercept the collective: keep the local gradient, gather all of them, compare with the reduction
    seen = {}
    real = train_step.allreduce_flat

    def spy(flat, ws):
        local_g = flat.clone()
        parts = [torch.empty_like(flat) for _ in range(ws)]
        dist.all_gather(parts, local_g)
        scale = real(flat, ws)
        seen["err"] = float((flat * scale - torch.stack(parts).mean(0)).abs().max())
        seen["mag"] = float(torch.stack(parts).mean(0).abs().max())
        return scale
    train_step.allreduce_flat = spy
    tr.step(*mine)
    train_step.allreduce_flat = real
    assert seen["err"] <= 1e-6 * max(seen["mag"], 1.0), seen
    # (1)+(3) weights identical across ranks
    def same_everywhere(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return all(bool((p == parts[0]).all()) for p in parts)
    assert same_everywhere(tr.flat), "weights diverged after the eager step"
    assert tr.capture(*mine), tr._capture_error
    tr.step_graph(*mine)
    tr.step_graph(*mine)
    torch.cuda.synchronize()
    assert same_everywhere(tr.flat), "weights diverged after graph replay"
    assert same_everywhere(tr.m) and same_everywhere(tr.v)
    print("DP_OK rank %d grad err %.3g (|g| max %.3g)" % (rank, seen["err"], seen["mag"]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
