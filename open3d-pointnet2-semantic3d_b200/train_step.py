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
    total = global_batch.shape[0]
    if world_size < 1 or not 0 <= rank < world_size or total % world_size != 0:
        raise ValueError("cannot shard %d clouds over %d ranks (rank %d): the batch must divide evenly"
                         % (total, world_size, rank))
    per = total // world_size
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
        tf_util.zero_arena.reset(point_cloud.device)  # one memset for every layer's fp64 accumulators
        try:
            pred, _ = model.get_model(point_cloud, True, self.num_class, self.params, bn_decay=bn_decay)
            if self.store.flat_params is None:   # first call created the variables: flatten, redo
                self._ensure_flat()
                tf_util.zero_arena.reset(point_cloud.device)
                pred, _ = model.get_model(point_cloud, True, self.num_class, self.params,
                                          bn_decay=bn_decay)
            self.store.zero_grad()
            loss = model.get_loss(pred, labels, smpw)
            loss.backward()
        finally:
            tf_util.zero_arena.disarm()
        return loss

    def step(self, point_cloud, labels, smpw):
        loss = self.forward_backward(point_cloud, labels, smpw)
        return self._apply_gradients(loss)

    def _apply_gradients(self, loss):
        gscale = allreduce_flat(self.grads, self.world_size)
        self.step_count += 1
        lr = get_learning_rate(self.step_count - 1, self.params)
        call("pn2_adam_step", self.flat.numel(), ptr(self.flat, F32), ptr(self.grads, F32),
             ptr(self.m, F32), ptr(self.v, F32), float(lr), 0.9, 0.999, 1e-8, self.step_count,
             float(gscale))
        return loss

    # ---- CUDA-graph mode: forward + loss + backward captured once, replayed per step -------------
    def capture(self, point_cloud, labels, smpw):
        """Capture forward+loss+backward of this batch shape into a CUDA graph (the ~370 launches
        of a step are otherwise CPU-launch bound).  The dropout mask stays fresh through a
        device-resident seed increment; schedule values that are baked in (bn_decay) trigger a
        re-capture when they change.  Returns False (and stays in eager mode) if capture fails."""
        from . import _ffi
        self._graph = None
        for mode in ("global", "thread_local"):
            try:
                self._static = [t.clone() for t in (point_cloud, labels, smpw)]
                self._seed_dev = torch.zeros(1, dtype=torch.int64, device=point_cloud.device)
                tf_util.set_dropout_seed_device(self._seed_dev)
                for _ in range(2):  # eager passes: every kernel / workspace / gradient buffer exists
                    self.forward_backward(*self._static)
                torch.cuda.synchronize()
                static = self._static

                def fb():
                    return self.forward_backward(*static)

                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fb()
                    fb()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = _ffi.launches
                with torch.cuda.graph(g, capture_error_mode=mode):
                    loss = fb()
                self._static_loss = loss
                self.launches_per_replay = _ffi.launches - n0
                g.replay()
                torch.cuda.synchronize()
                self._graph = g
                self._graph_bn_decay = get_bn_decay(self.step_count, self.params)
                self._capture_error = None
                return True
            except Exception as e:  # noqa: BLE001 - any capture failure means eager mode
                import traceback
                self._graph = None
                self._capture_error = "[%s] " % mode + repr(e)[:160] + " | " + " <- ".join(
                    l.strip() for l in traceback.format_exc().splitlines()
                    if l.strip().startswith("File"))[-700:]
                try:
                    torch.cuda.synchronize()
                except Exception:  # noqa: BLE001
                    pass
        tf_util.set_dropout_seed_device(None)
        return False

    def step_graph(self, point_cloud, labels, smpw):
        if getattr(self, "_graph", None) is None:
            return self.step(point_cloud, labels, smpw)
        if get_bn_decay(self.step_count, self.params) != self._graph_bn_decay:
            if not self.capture(point_cloud, labels, smpw):
                return self.step(point_cloud, labels, smpw)
        for dst, src in zip(self._static, (point_cloud, labels, smpw)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._seed_dev.add_(1)
        self._graph.replay()
        return self._apply_gradients(self._static_loss)
