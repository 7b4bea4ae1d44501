"""EXPERIMENTAL (never run on a GPU yet): geometry one batch ahead.

Farthest point sampling, gather_point, ball query, 3-NN and the inverse-distance weights depend only on
the point coordinates, never on the weights -- yet they sit at the head of the training step's critical
path (~1.0 ms of 4.87 ms in round 1; FPS alone keeps 16 of 148 SMs busy for 0.57 ms).  This module runs
that weight-independent part ("geometry") for batch i+1 on a side stream while the dense part of batch i
(grouping, shared MLPs, loss, backward) runs on the main stream:

    main :  dense(i) -- reads geometry S --------------->  wait side, S <- N, Adam
    side :  geometry(i+1) -> N  (same kernels, same order as model.get_model)

Nothing is cached across steps: K steps compute K batches of geometry (plus one to fill the pipe), and the
values the dense stage consumes are the ones the plain Trainer would compute, so losses and updates are the
same (tests/test_experimental_gpu.py compares them step by step).

How it plugs in without touching the verified default path: `PrefetchTrainer` is a subclass of `Trainer`;
while it is alive, the five geometry ops that `util/pointnet_util.py` looks up in its own module namespace
are wrapped by `_Replay`, which hands back the precomputed tensors in the order the model asks for them
(and checks that order); with no tape set the wrappers call straight through.
"""
import torch

from .tf_ops import tf_grouping, tf_interpolate, tf_sampling
from .train_step import Trainer, get_bn_decay
from .util import pointnet_util

GEOM_OPS = ("farthest_point_sample", "gather_point", "query_ball_point", "three_nn", "fp_weights")


def compute_geometry(point_cloud, hp):
    """The weight-independent ops of model.get_model (model.py:36-129), in its call order.
    Returns the tape [(op name, result)], results exactly as the ops return them."""
    with torch.no_grad():
        xyz = point_cloud[:, :, 0:3].contiguous() if hp["use_color"] else point_cloud
        tape, levels = [], [xyz]
        for l in (1, 2, 3, 4):
            idx = tf_sampling.farthest_point_sample(hp["l%d_npoint" % l], levels[-1])
            tape.append(("farthest_point_sample", idx))
            new_xyz = tf_sampling.gather_point(levels[-1], idx)
            tape.append(("gather_point", new_xyz))
            tape.append(("query_ball_point", tf_grouping.query_ball_point(
                hp["l%d_radius" % l], hp["l%d_nsample" % l], levels[-1], new_xyz)))
            levels.append(new_xyz)
        for lo in (3, 2, 1, 0):
            dist, idx = tf_interpolate.three_nn(levels[lo], levels[lo + 1])
            tape.append(("three_nn", (dist, idx)))
            tape.append(("fp_weights", _REAL["fp_weights"](dist)))
    return tape


def _tensors(tape):
    out = []
    for _, v in tape:
        out.extend(v if isinstance(v, tuple) else (v,))
    return out


def _clone_tape(tape):
    return [(n, tuple(t.clone() for t in v) if isinstance(v, tuple) else v.clone()) for n, v in tape]


class _Replay:
    """Stands in for the geometry ops inside pointnet_util while a tape is set."""

    def __init__(self):
        self.tape, self.pos = None, 0

    def wrap(self, name, real):
        def op(*args, **kwargs):
            if self.tape is None:
                return real(*args, **kwargs)
            if self.pos == len(self.tape):  # a new forward pass over the same batch
                self.pos = 0
            expected, value = self.tape[self.pos]
            if expected != name:
                raise RuntimeError("geometry tape out of order: the model called %s where the tape holds %s "
                                   "(only model.get_model's SSG sequence can be prefetched)" % (name, expected))
            self.pos += 1
            return value
        op.__name__ = name
        return op


_REAL = {name: getattr(pointnet_util, name) for name in GEOM_OPS}
_replay = _Replay()
_installed = [0]


def _install():
    if _installed[0] == 0:
        for name in GEOM_OPS:
            setattr(pointnet_util, name, _replay.wrap(name, _REAL[name]))
    _installed[0] += 1


def _uninstall():
    _installed[0] -= 1
    if _installed[0] == 0:
        for name in GEOM_OPS:
            setattr(pointnet_util, name, _REAL[name])
        _replay.tape, _replay.pos = None, 0


class PrefetchTrainer(Trainer):
    """Trainer whose geometry runs one batch ahead on a side stream.

        tr = PrefetchTrainer(hp, num_class)
        tr.prime(first_batch_pc)                       # fills the pipe: geometry of batch 0
        loss = tr.step_prefetch(pc, labels, smpw, next_pc)        # eager launches
        tr.capture_prefetch(pc, labels, smpw); loss = tr.step_graph_prefetch(pc, labels, smpw, next_pc)
        tr.close()
    """

    def __init__(self, params, num_class, device="cuda", seed=0, world_size=1):
        super().__init__(params, num_class, device=device, seed=seed, world_size=world_size)
        self._side = torch.cuda.Stream(device=device)
        self._S = None        # geometry of the batch the next dense stage consumes
        self._N = None        # geometry being produced for the batch after that
        self._g_geom = None
        _install()

    def close(self):
        _uninstall()

    # ---- the dense stage consumes the current tape ----------------------------------------------------
    def forward_backward(self, point_cloud, labels, smpw):
        if self._S is None:
            raise RuntimeError("PrefetchTrainer: call prime(point_cloud) before the first step")
        _replay.tape, _replay.pos = self._S, 0
        try:
            return super().forward_backward(point_cloud, labels, smpw)
        finally:
            _replay.tape = None

    def prime(self, point_cloud):
        """Pipeline fill: geometry of the first batch, on the main stream."""
        self._S = _clone_tape(compute_geometry(point_cloud, self.params))

    # ---- eager mode -----------------------------------------------------------------------------------
    def step_prefetch(self, point_cloud, labels, smpw, next_point_cloud):
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            nxt = compute_geometry(next_point_cloud, self.params)
        loss = self.forward_backward(point_cloud, labels, smpw)
        main.wait_stream(self._side)
        for t in _tensors(nxt):
            t.record_stream(main)  # produced on the side stream, consumed on the main one
        self._S = nxt
        return self._apply_gradients(loss)

    # ---- CUDA-graph mode: one graph for the dense stage (reads S), one for the geometry (writes N) ----------
    def capture_prefetch(self, point_cloud, labels, smpw):
        self.prime(point_cloud)
        if not self.capture(point_cloud, labels, smpw):  # the parent's capture, through forward_backward above
            return False
        try:
            self._next_pc = point_cloud.clone()
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                for _ in range(2):
                    compute_geometry(self._next_pc, self.params)
            torch.cuda.synchronize()
            from . import _ffi
            g = torch.cuda.CUDAGraph()
            n0 = _ffi.launches
            with torch.cuda.graph(g, stream=self._side):
                self._N = compute_geometry(self._next_pc, self.params)
            self.geom_launches_per_replay = _ffi.launches - n0
            torch.cuda.synchronize()
            self._g_geom = g
            return True
        except Exception as e:  # noqa: BLE001 - stay usable in eager mode
            self._g_geom, self._graph = None, None
            self._capture_error = "geometry graph: " + repr(e)[:300]
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
            return False

    def step_graph_prefetch(self, point_cloud, labels, smpw, next_point_cloud):
        if getattr(self, "_graph", None) is None or self._g_geom is None:
            return self.step_prefetch(point_cloud, labels, smpw, next_point_cloud)
        if get_bn_decay(self.step_count, self.params) != self._graph_bn_decay:
            keep = _clone_tape(self._S)  # the geometry of THIS batch survives the re-capture
            if not self.capture_prefetch(point_cloud, labels, smpw):
                self._S = keep
                return self.step_prefetch(point_cloud, labels, smpw, next_point_cloud)
            for d, s in zip(_tensors(self._S), _tensors(keep)):
                d.copy_(s)
        main = torch.cuda.current_stream()
        for dst, src in zip(self._static, (point_cloud, labels, smpw)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._side.wait_stream(main)          # previous step's S <- N copy is ordered before the next N
        with torch.cuda.stream(self._side):
            if self._next_pc.data_ptr() != next_point_cloud.data_ptr():
                self._next_pc.copy_(next_point_cloud, non_blocking=True)
            self._g_geom.replay()
        self._seed_dev.add_(1)
        self._graph.replay()                  # dense stage of this batch, concurrent with the side stream
        main.wait_stream(self._side)
        for d, s in zip(_tensors(self._S), _tensors(self._N)):
            d.copy_(s, non_blocking=True)     # a few MB of indices: the next batch's geometry becomes current
        return self._apply_gradients(self._static_loss)
