"""CPU: the reference's OWN model.py, unmodified, imported from /root/reference and run on this engine's
API surface through the tensorflow shim (compat/tensorflow.py) -- SURVEY.md section 7 hard part 7.

The device kernels cannot run here, so the four functions model.py calls into (pointnet_sa_module,
pointnet_fp_module, tf_util.conv1d, tf_util.dropout) and the loss are replaced by recorders that (1) bind every
call against the REAL function's signature (a wrong keyword or a missing argument fails), (2) log the normalised
arguments and tensor shapes, (3) return CPU tensors of the right shapes.  The same recorders then run this
package's model.py: both must make exactly the same calls in the same order -- which makes the package's
model.py (the one the GPU parity tests cover) call-for-call equivalent to the reference file.
Skipped where /root/reference is absent (the GPU box)."""
import importlib.util
import inspect
import os
import sys

import numpy as np
import pytest

REF_MODEL = "/root/reference/model.py"


def _norm(v):
    import torch
    if isinstance(v, torch.Tensor):
        return ("tensor", tuple(v.shape), str(v.dtype))
    if isinstance(v, (list, tuple)):
        return tuple(_norm(x) for x in v)
    return v


@pytest.mark.skipif(not os.path.exists(REF_MODEL), reason="reference tree not present")
def test_reference_model_py_runs_unmodified_and_matches_our_model(monkeypatch):
    import torch
    import pn2_b200
    from pn2_b200 import model as ours
    from pn2_b200.compat import tensorflow as tfs
    from pn2_b200.util import pointnet_util as pu, tf_util

    log = []

    def recorder(name, real, make_result):
        sig = inspect.signature(real)

        def fn(*a, **kw):
            bound = sig.bind(*a, **kw)          # TypeError on a call the real function would reject
            bound.apply_defaults()
            log.append((name, tuple((k, _norm(v)) for k, v in bound.arguments.items())))
            return make_result(bound.arguments)
        return fn

    def sa_result(a):
        b = a["xyz"].shape[0]
        return (torch.zeros(b, a["npoint"], 3), torch.zeros(b, a["npoint"], a["mlp"][-1]),
                torch.zeros(b, a["npoint"], a["nsample"], dtype=torch.int32))

    fakes = {
        "pointnet_sa_module": recorder("pointnet_sa_module", pu.pointnet_sa_module, sa_result),
        "pointnet_fp_module": recorder("pointnet_fp_module", pu.pointnet_fp_module,
                                       lambda a: torch.zeros(a["xyz1"].shape[0], a["xyz1"].shape[1], a["mlp"][-1])),
        "conv1d": recorder("conv1d", tf_util.conv1d,
                           lambda a: torch.zeros(*a["inputs"].shape[:-1], a["num_output_channels"])),
        "dropout": recorder("dropout", tf_util.dropout, lambda a: a["inputs"]),
        "get_loss": recorder("get_loss", ours.get_loss, lambda a: torch.zeros(())),
    }
    monkeypatch.setattr(pu, "pointnet_sa_module", fakes["pointnet_sa_module"])
    monkeypatch.setattr(pu, "pointnet_fp_module", fakes["pointnet_fp_module"])
    monkeypatch.setattr(tf_util, "conv1d", fakes["conv1d"])
    monkeypatch.setattr(tf_util, "dropout", fakes["dropout"])
    monkeypatch.setattr(ours, "pointnet_sa_module", fakes["pointnet_sa_module"])  # bound at import time
    monkeypatch.setattr(ours, "pointnet_fp_module", fakes["pointnet_fp_module"])
    monkeypatch.setattr(ours, "get_loss", fakes["get_loss"])

    before = set(sys.modules)
    try:
        assert tfs.install(), "a real tensorflow is importable here?"
        spec = importlib.util.spec_from_file_location("reference_model_py", REF_MODEL)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)            # `import tensorflow`, `import util.tf_util`, ... resolve to this package
        hp = {"use_color": 1, "l1_npoint": 32, "l1_radius": 0.5, "l1_nsample": 8, "l2_npoint": 16, "l2_radius": 1.0,
              "l2_nsample": 8, "l3_npoint": 8, "l3_radius": 2.0, "l3_nsample": 4, "l4_npoint": 4, "l4_radius": 4.0,
              "l4_nsample": 4}
        pc = torch.as_tensor(np.random.RandomState(0).random_sample((2, 64, 6)).astype(np.float32))
        labels = torch.zeros(2, 64, dtype=torch.int32)
        smpw = torch.ones(2, 64)

        # placeholders: dtype/shape records (model.py:12-19)
        pls = ref.get_placeholders(64, hp)
        assert [p.shape for p in pls] == [(None, 64, 6), (None, 64), (None, 64)]
        assert [p.dtype for p in pls] == [torch.float32, torch.int32, torch.float32]
        assert [tuple(p.shape) for p in ours.get_placeholders(64, hp)] == [p.shape for p in pls]

        pred, end_points = ref.get_model(pc, True, 9, hp, bn_decay=0.5)
        ref.get_loss(pred, labels, smpw, end_points)
        ref_log = list(log)
        assert tuple(pred.shape) == (2, 64, 9) and set(end_points) == {"l0_xyz", "feats"}
        assert torch.equal(end_points["l0_xyz"], pc[:, :, :3])                       # tf.slice
        assert "classify loss" in tfs.summary.values and len(tfs.get_collection("losses")) >= 1

        del log[:]
        pred2, end_points2 = ours.get_model(pc, True, 9, hp, bn_decay=0.5)
        ours.get_loss(pred2, labels, smpw, end_points2)
        assert len(ref_log) == 4 + 4 + 2 + 1 + 1                                      # SA, FP, conv1d, dropout, loss
        assert [c[0] for c in ref_log] == [c[0] for c in log]
        strip = lambda c: (c[0], tuple(kv for kv in c[1] if kv[0] != "end_points"))  # noqa: E731 (unused by both)
        for a, b in zip(ref_log, log):
            assert strip(a) == strip(b), (a, b)

        # use_color = 0 (semantic_no_color.json): no slicing, points=None into layer1
        del log[:]
        ref.get_model(pc[:, :, :3].contiguous(), False, 9, dict(hp, use_color=0))
        first = dict(log[0][1])
        assert first["points"] is None and first["xyz"] == ("tensor", (2, 64, 3), "torch.float32")
    finally:
        for k in set(sys.modules) - before:
            if k == "tensorflow" or k.split(".")[0] in ("tf_ops", "util", "reference_model_py"):
                sys.modules.pop(k, None)
