"""One training step of the reference's train.py on the sm_100a engine.

  get_learning_rate / get_bn_decay                       train.py:80-119
  Trainer.step: forward, loss, backward, Adam            train.py:387-388, 225-244
  data parallelism: batch sharded across ranks, ONE NCCL all-reduce over the flat gradient
  buffer per step; BatchNorm statistics stay per replica.
"""
import math

import torch
import torch.distributed as dist

from . import model
from ._ffi import F32, call, ptr
from .util import tf_util


def get_learning_rate(step, params):
    """tf.train.exponential_decay(staircase) clipped at 1e-5 (train.py:80-97)."""
    lr = params["learning_rate"] * params["learning_rate_decay_rate"] ** (
        (step * params["batch_size"]) // params["decay_step"])
    return max(lr, 0.00001)


def get_bn_decay(step, params):
    """min(clip, 1 - init*rate^floor(step*B/decay_step)) (train.py:100-119)."""
    mom = params["bn_init_decay"] * params["bn_decay_decay_rate"] ** (
        (step * params["batch_size"]) // int(params["decay_step"]))
    return min(params["bn_decay_clip"], 1 - mom)


def shard_batch(global_batch, rank, world_size):
    """Contiguous shard of the leading (cloud) dimension owned by ``rank``: the path is
    independent per cloud, so data parallelism partitions clouds with no data-path exchange."""
    per = global_batch.shape[0] // world_size
    return global_batch[rank * per:(rank + 1) * per]


def allreduce_flat(flat_grads, world_size):
    """The single collective of a step: SUM all-reduce of the flat gradient buffer.  Returns the
    scale (1/world_size) the Adam kernel folds into its gradient read."""
    if world_size > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / world_size


class Trainer:
    """Owns the variables, Adam moments and the step counter for one replica."""

    def __init__(self, params, num_class, device="cuda", seed=0, world_size=1):
        self.params, self.num_class, self.world_size = params, num_class, world_size
        self.store = tf_util.set_default_store(tf_util.VariableStore(device=device, seed=seed))
        self.step_count = 0
        self.flat = self.grads = self.m = self.v = None

    def _ensure_flat(self):
        if self.store.flat_params is None:
            self.flat, self.grads = self.store.flatten()
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)
            if self.world_size > 1:  # replicas start from rank 0's weights
                dist.broadcast(self.flat, src=0)

    def forward_backward(self, point_cloud, labels, smpw):
        bn_decay = get_bn_decay(self.step_count, self.params)
        tf_util.set_default_store(self.store)
        pred, _ = model.get_model(point_cloud, True, self.num_class, self.params, bn_decay=bn_decay)
        if self.store.flat_params is None:   # first call created the variables: flatten, redo
            self._ensure_flat()
            pred, _ = model.get_model(point_cloud, True, self.num_class, self.params,
                                      bn_decay=bn_decay)
        self.store.zero_grad()
        loss = model.get_loss(pred, labels, smpw)
        loss.backward()
        return loss

    def step(self, point_cloud, labels, smpw):
        loss = self.forward_backward(point_cloud, labels, smpw)
        gscale = allreduce_flat(self.grads, self.world_size)
        self.step_count += 1
        lr = get_learning_rate(self.step_count - 1, self.params)
        call("pn2_adam_step", self.flat.numel(), ptr(self.flat, F32), ptr(self.grads, F32),
             ptr(self.m, F32), ptr(self.v, F32), float(lr), 0.9, 0.999, 1e-8, self.step_count,
             float(gscale))
        return loss
