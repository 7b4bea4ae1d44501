"""CPU: host-side logic -- schedules, variable store / scoping, argument validation."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PARAMS = {"batch_size": 16, "learning_rate": 0.001, "decay_step": 200000,
          "learning_rate_decay_rate": 0.7, "bn_init_decay": 0.5, "bn_decay_decay_rate": 0.5,
          "bn_decay_clip": 0.99}


def test_schedules_follow_train_py():
    """train.py:80-119: staircase exponential decays, lr floor 1e-5, bn_decay clip 0.99."""
    from pn2_b200.train_step import get_bn_decay, get_learning_rate
    assert get_learning_rate(0, PARAMS) == 0.001
    assert get_learning_rate(12499, PARAMS) == 0.001          # 12499*16 < 200000
    assert abs(get_learning_rate(12500, PARAMS) - 0.0007) < 1e-12
    assert get_learning_rate(10 ** 7, PARAMS) == 0.00001
    assert get_bn_decay(0, PARAMS) == 0.5
    assert get_bn_decay(12500, PARAMS) == 0.75
    assert get_bn_decay(10 ** 7, PARAMS) == 0.99


def test_variable_store_scoping_and_flatten():
    import torch
    from pn2_b200.util import tf_util
    st = tf_util.set_default_store(tf_util.VariableStore(device="cpu", seed=3))
    with tf_util.variable_scope("layer1"):
        L = tf_util.make_layer("conv0", 6, 32, True, tf_util.relu)
    assert L.w.name == "layer1/conv0/weights" and tuple(L.w.data.shape) == (1, 1, 6, 32)
    assert set(st.vars) == {"layer1/conv0/weights", "layer1/conv0/biases", "layer1/conv0/bn/gamma",
                            "layer1/conv0/bn/beta", "layer1/conv0/bn/moving_mean",
                            "layer1/conv0/bn/moving_variance"}
    lim = np.sqrt(6.0 / (6 + 32))
    assert float(L.w.data.abs().max()) <= lim and float(L.b.data.abs().max()) == 0.0
    assert float(L.gamma.data.min()) == 1.0 and float(L.mv.data.min()) == 1.0
    # same scope again returns the same variables (reuse), shape mismatch is an error
    with tf_util.variable_scope("layer1"):
        L2 = tf_util.make_layer("conv0", 6, 32, True, tf_util.relu)
        assert L2.w is L.w
        with pytest.raises(ValueError):
            tf_util.make_layer("conv0", 7, 32, True, tf_util.relu)
    flat, grads = st.flatten()
    assert flat.numel() == 6 * 32 + 32 * 3 and grads.numel() == flat.numel()
    L.w.grad += 1.0
    assert float(grads.sum()) == 6 * 32
    st.zero_grad()
    assert float(grads.abs().sum()) == 0.0
    # data are views of the flat buffer
    flat.fill_(2.0)
    assert float(L.w.data.mean()) == 2.0 and not L.mm.trainable


def test_python_validation_messages_match_reference():
    """OP_REQUIRES texts of tf_sampling.cpp / tf_grouping.cpp / tf_interpolate.cpp."""
    import torch
    from pn2_b200.tf_ops import tf_grouping, tf_interpolate, tf_sampling
    x = torch.rand(1, 8, 3)
    with pytest.raises(ValueError, match="FarthestPointSample expects positive npoint"):
        tf_sampling.farthest_point_sample(0, x)
    with pytest.raises(ValueError, match=r"FarthestPointSample expects \(batch_size,num_points,3\)"):
        tf_sampling.farthest_point_sample(4, x[..., :2])
    with pytest.raises(ValueError, match="QueryBallPoint expects positive radius"):
        tf_grouping.query_ball_point(-1.0, 4, x, x)
    with pytest.raises(ValueError, match="QueryBallPoint expects positive nsample"):
        tf_grouping.query_ball_point(1.0, 0, x, x)
    with pytest.raises(ValueError, match="GroupPoint expects"):
        tf_grouping.group_point(x[0], torch.zeros(1, 2, 2, dtype=torch.int32))
    with pytest.raises(ValueError, match="ThreeNN expects"):
        tf_interpolate.three_nn(x[..., :2], x)
    with pytest.raises(ValueError, match="ThreeInterpolate expects"):
        tf_interpolate.three_interpolate(x, torch.zeros(1, 4, 2, dtype=torch.int32), torch.zeros(1, 4, 3))
    with pytest.raises(ValueError, match="SelectionSort expects positive k"):
        tf_grouping.select_top_k(0, torch.rand(1, 2, 3))
    with pytest.raises(ValueError, match=r"ProbSample expects \(batch_size,num_choices\) inp shape"):
        tf_sampling.prob_sample(torch.rand(2, 5, 1), torch.rand(2, 3))
    with pytest.raises(ValueError, match=r"ProbSample expects \(batch_size,num_points\) inpr shape"):
        tf_sampling.prob_sample(torch.rand(2, 5), torch.rand(3, 3))
    lab = torch.zeros(8, dtype=torch.int32)
    with pytest.raises(ValueError, match="sparse_points must be"):
        tf_interpolate.interpolate_label_with_color(x[0][:, :2], lab, x[0], 3)
    with pytest.raises(ValueError, match="sparse_labels must be"):
        tf_interpolate.interpolate_label_with_color(x[0], lab[:5], x[0], 3)
    with pytest.raises(ValueError, match="dense_points must be"):
        tf_interpolate.interpolate_label_with_color(x[0], lab, x, 3)
    with pytest.raises(ValueError, match="knn must be an int scalar"):
        tf_interpolate.interpolate_label_with_color(x[0], lab, x[0], 2.5)


def test_reference_aliases():
    import sys
    import pn2_b200
    pn2_b200.install_reference_aliases()
    from tf_ops.tf_sampling import prob_sample, farthest_point_sample, gather_point  # noqa: F401
    from tf_ops.tf_grouping import query_ball_point, group_point, knn_point  # noqa: F401
    from tf_ops.tf_interpolate import three_nn, three_interpolate  # noqa: F401
    from tf_ops.tf_interpolate import interpolate_label_with_color  # noqa: F401
    from util.pointnet_util import (pointnet_sa_module, pointnet_sa_module_msg,  # noqa: F401
                                    pointnet_fp_module, sample_and_group, sample_and_group_all)
    from util import tf_util
    assert hasattr(tf_util, "conv2d") and hasattr(tf_util, "conv1d") and hasattr(tf_util, "dropout")
    from predict import Predictor  # noqa: F401
    import model
    assert hasattr(model, "get_model") and hasattr(model, "get_loss")
    for k in ["tf_ops", "tf_ops.tf_sampling", "tf_ops.tf_grouping", "tf_ops.tf_interpolate", "util",
              "util.tf_util", "util.pointnet_util", "model", "predict"]:
        sys.modules.pop(k, None)


def test_layer_signatures_match_reference():
    """Positional/keyword orders of pointnet_util.py:98-116, 219-232, 285-287 and model.py."""
    import inspect
    from pn2_b200.util import pointnet_util as pu
    from pn2_b200 import model
    assert list(inspect.signature(pu.pointnet_sa_module).parameters) == [
        "xyz", "points", "npoint", "radius", "nsample", "mlp", "mlp2", "group_all", "is_training",
        "bn_decay", "scope", "bn", "pooling", "knn", "use_xyz", "use_nchw"]
    assert list(inspect.signature(pu.pointnet_sa_module_msg).parameters) == [
        "xyz", "points", "npoint", "radius_list", "nsample_list", "mlp_list", "is_training",
        "bn_decay", "scope", "bn", "use_xyz", "use_nchw"]
    assert list(inspect.signature(pu.pointnet_fp_module).parameters) == [
        "xyz1", "xyz2", "points1", "points2", "mlp", "is_training", "bn_decay", "scope", "bn"]
    assert list(inspect.signature(pu.sample_and_group).parameters) == [
        "npoint", "radius", "nsample", "xyz", "points", "knn", "use_xyz"]
    assert list(inspect.signature(model.get_model).parameters) == [
        "point_cloud", "is_training", "num_class", "hyperparams", "bn_decay"]
    ph = model.get_placeholders(8192, {"use_color": 1})
    assert ph[0].shape == (None, 8192, 6) and ph[1].shape == (None, 8192)


def test_shard_batch_is_even_and_contiguous():
    import torch
    from pn2_b200.train_step import shard_batch
    g = torch.arange(8 * 3).reshape(8, 3)
    parts = [shard_batch(g, r, 4) for r in range(4)]
    assert all(p.shape == (2, 3) for p in parts) and torch.equal(torch.cat(parts), g)
    with pytest.raises(ValueError, match="divide evenly"):
        shard_batch(g, 0, 3)
    with pytest.raises(ValueError):
        shard_batch(g, 4, 4)


def test_predictor_host_logic(monkeypatch, tmp_path):
    """predict.py:15-105 mirror: checkpoint dict / npz loading, eval-mode call, arg-max, label transfer
    plumbing.  The device work is replaced by stubs here (the ops themselves are GPU-tested)."""
    import torch
    from pn2_b200 import predict
    hp = {"use_color": 1}
    ck = {"layer1/conv0/weights": np.ones((1, 1, 6, 4), np.float32),
          "layer1/conv0/bn/moving_mean": np.zeros(4, np.float32)}
    np.savez(tmp_path / "ck.npz", **ck)
    seen = {}

    def fake_get_model(x, is_training, num_class, hyperparams, bn_decay=None):
        seen["args"] = (tuple(x.shape), is_training, num_class, torch.is_grad_enabled())
        logits = torch.zeros(x.shape[0], x.shape[1], num_class)
        logits[:, :, 3] = 1.0
        logits[0, 0, 5] = 2.0
        return logits, {}

    def fake_vote(sp, sl, dp, knn):
        seen["vote"] = (sp.dtype, sl.dtype, tuple(dp.shape), knn)
        return torch.full((dp.shape[0],), 7, dtype=torch.int32), torch.zeros((dp.shape[0], 3), dtype=torch.uint8)

    monkeypatch.setattr(predict.model, "get_model", fake_get_model)
    monkeypatch.setattr(predict, "interpolate_label_with_color", fake_vote)
    for src in (ck, str(tmp_path / "ck.npz")):
        p = predict.Predictor(src, 9, hp, device="cpu")
        assert set(p.store.vars) == set(ck) and not p.store.vars["layer1/conv0/bn/moving_mean"].trainable
        labels = p.predict(np.zeros((2, 16, 6), np.float32))
        assert labels.shape == (2, 16) and labels[0, 0] == 5 and labels[1, 3] == 3
        assert seen["args"] == ((2, 16, 6), False, 9, False)
    with pytest.raises(ValueError, match="batch_data must be"):
        p.predict(np.zeros((2, 16, 3), np.float32))
    with pytest.raises(ValueError, match="checkpoint"):
        predict.Predictor({}, 9, hp, device="cpu")
    dl, dc = p.interpolate_labels(np.zeros((5, 3)), np.zeros(5, np.int64), np.zeros((11, 3)))
    assert dl.tolist() == [7] * 11 and dc.shape == (11, 3)
    assert seen["vote"] == (torch.float32, torch.int32, (11, 3), 3)


def test_bench_algorithmic_work_matches_survey():
    """bench.py's roofline arithmetic against the figures SURVEY.md 8(d) states for config 2 (B=16):
    39.6 GFLOP of GEMM forward over 23 conv layers, FPS streaming model 2.68 GB for SA1, GEMM bytes
    4*M*(K+N) per call."""
    import bench
    calls = bench.linear_calls(16)
    assert len(calls) == 23
    assert calls[0] == (16 * 1024 * 32, 6, 32) and calls[-1] == (16 * 8192, 128, 9)
    assert (16 * 8192, 128 + 3, 128) in calls and (16 * 64, 512 + 256, 256) in calls  # FP4 / FP1 first layers
    assert abs(bench.gemm_flops(16) / 1e9 - 39.6) < 0.05
    fwd, dgrad, wgrad = bench.gemm_bytes(16)
    assert fwd == sum(4 * m * (k + n) for m, k, n in calls) == wgrad
    assert dgrad == fwd - 4 * calls[0][0] * (calls[0][1] + calls[0][2])  # no dgrad into the network input
    fps = bench.fps_stream_bytes(16)
    assert fps[0] == 16 * 1023 * 8192 * 20 and abs(fps[0] / 1e9 - 2.68) < 0.01
    pc, labels, smpw = bench.make_batch(2, 64, 100)
    assert pc.shape == (2, 64, 6) and pc.dtype == np.float32 and labels.min() >= 1 and labels.max() <= 8
    assert pc[..., 0].min() >= -5 and pc[..., 0].max() <= 5 and pc[..., 2].min() >= 0 and pc[..., 2].max() <= 5


def test_load_state_dict_is_strict_about_shapes_and_foreign_keys():
    """a wrong-shaped checkpoint array must not be broadcast into a variable, optimizer slots of
    a full TF checkpoint must not become variables, and a strict store never initialises a missing one."""
    import torch
    from pn2_b200.util import tf_util
    st = tf_util.VariableStore(device="cpu", seed=0)
    tf_util.set_default_store(st)
    L = tf_util.make_layer("layer1/conv0", 6, 32, True, tf_util.relu)
    w = np.arange(6 * 32, dtype=np.float32).reshape(6, 32)
    skipped = st.load_state_dict({"layer1/conv0/weights": w,                    # [k,n] into [1,1,k,n]: allowed
                                  "layer1/conv0/weights/Adam": np.zeros((1, 1, 6, 32), np.float32),
                                  "beta1_power": np.float32(0.9),
                                  "layer9/conv0/bn/moving_mean": np.zeros(4, np.float32)})
    assert sorted(skipped) == ["beta1_power", "layer1/conv0/weights/Adam"]
    assert torch.equal(L.w.data.reshape(6, 32), torch.as_tensor(w))
    assert "layer9/conv0/bn/moving_mean" in st.vars and not st.vars["layer9/conv0/bn/moving_mean"].trainable
    assert "beta1_power" not in st.vars and "layer1/conv0/weights/Adam" not in st.vars
    with pytest.raises(ValueError, match="shape"):
        st.load_state_dict({"layer1/conv0/bn/gamma": np.ones((1, 1, 6, 32), np.float32)})  # would broadcast
    with pytest.raises(ValueError, match="shape"):
        st.load_state_dict({"layer1/conv0/weights": np.zeros((32, 6), np.float32)})        # transposed
    st.strict = True
    tf_util.make_layer("layer1/conv0", 6, 32, True, tf_util.relu)  # exists: fine
    with pytest.raises(KeyError, match="not in the loaded checkpoint"):
        tf_util.make_layer("layer1/conv1", 32, 32, True, tf_util.relu)
    # a [k,n] kernel loaded BEFORE the layer exists adopts the layer's rank on first use
    st2 = tf_util.set_default_store(tf_util.VariableStore(device="cpu", seed=0))
    st2.load_state_dict({"fc1/weights": np.zeros((128, 128), np.float32)})
    L2 = tf_util.make_layer("fc1", 128, 128, False, None, kernel_rank=3)
    assert tuple(L2.w.data.shape) == (1, 128, 128)


def test_every_kernel_waits_for_its_predecessor():
    """Programmatic dependent launch (csrc/pn2_common.cuh): every kernel of the library is launched with the
    programmatic-serialization attribute, so every `__global__` function has to execute griddepcontrol.wait
    (pdl_enter / pdl_wait) before it touches global memory -- a kernel without it would start
    reading while its predecessor is still writing.  Static check over the sources: the wait is there, it comes
    before the first global-memory access idiom, and no launch bypasses launch_k."""
    import glob
    import re
    src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           "open3d-pointnet2-semantic3d_b200", "csrc")
    kernels = 0
    for path in sorted(glob.glob(os.path.join(src_dir, "*.cu")) + glob.glob(os.path.join(src_dir, "*.cuh"))):
        text = open(path).read()
        assert "<<<" not in re.sub(r"//[^\n]*", "", text), "%s launches a kernel outside launch_k" % path
        for m in re.finditer(r"__global__", text):
            i, depth, seen = m.end(), 0, False
            while True:  # body start: the first '{' outside parentheses after the parameter list
                c = text[i]
                if c == "(":
                    depth, seen = depth + 1, True
                elif c == ")":
                    depth -= 1
                elif c == "{" and depth == 0 and seen:
                    break
                elif c == ";" and depth == 0:
                    i = -1
                    break
                i += 1
            if i < 0:
                continue  # a declaration
            j, depth = i, 0
            while True:
                depth += {"{": 1, "}": -1}.get(text[j], 0)
                if depth == 0:
                    break
                j += 1
            body = text[i:j]
            name = re.findall(r"(\w+)\s*\($", text[m.end():i].split("(")[0] + "(")  # best effort, for the message
            w = re.search(r"pdl_enter\(\)|pdl_wait\(\)", body)
            assert w, "%s: kernel %s never waits for its predecessor" % (os.path.basename(path), name)
            before = body[:w.start()]
            for idiom in ("__ldg", "__ldcg", "atomicAdd", "cp.async", "ld.global", "st.global", "red.global"):
                assert idiom not in before, "%s: %s before the dependency wait" % (os.path.basename(path), idiom)
            kernels += 1
    assert kernels >= 60


def test_geometry_tape_records_and_replays_in_model_order(monkeypatch):
    """model.get_geometry runs the weight-independent ops of get_model (FPS, gather, ball query per SA layer; 3-NN and
    weights per FP layer) in the model's order and files the results in a GeometryTape; under replay_geometry the
    layers' two geometry hooks hand the stored tensors back WITHOUT launching anything, insist on the recorded order and
    arguments, and compute again inside replay_geometry(None).  Device ops faked on the CPU."""
    import torch
    import pn2_b200  # noqa: F401
    from pn2_b200 import model
    from pn2_b200.util import pointnet_util as pu
    calls = []

    def fps(npoint, xyz):
        calls.append(("fps", npoint, tuple(xyz.shape)))
        return torch.zeros(xyz.shape[0], npoint, dtype=torch.int32)

    def gather(xyz, idx):
        calls.append(("gather", tuple(xyz.shape), tuple(idx.shape)))
        return torch.full((xyz.shape[0], idx.shape[1], 3), float(len(calls)))

    def ball(radius, nsample, xyz, new_xyz):
        calls.append(("ball", radius, nsample, tuple(xyz.shape), tuple(new_xyz.shape)))
        return torch.full((xyz.shape[0], new_xyz.shape[1], nsample), len(calls), dtype=torch.int32), None

    def three_nn(xyz1, xyz2):
        calls.append(("three_nn", tuple(xyz1.shape), tuple(xyz2.shape)))
        return torch.ones(xyz1.shape[0], xyz1.shape[1], 3), torch.full((xyz1.shape[0], xyz1.shape[1], 3), len(calls), dtype=torch.int32)

    def weights(dist):
        calls.append(("weights", tuple(dist.shape)))
        return dist / 3

    for name, fn in (("farthest_point_sample", fps), ("gather_point", gather), ("query_ball_point", ball),
                     ("three_nn", three_nn), ("fp_weights", weights)):
        monkeypatch.setattr(pu, name, fn)
    hp = {"use_color": 1, "l1_npoint": 32, "l1_radius": 0.5, "l1_nsample": 8, "l2_npoint": 16, "l2_radius": 1.0,
          "l2_nsample": 8, "l3_npoint": 8, "l3_radius": 2.0, "l3_nsample": 4, "l4_npoint": 4, "l4_radius": 4.0,
          "l4_nsample": 4}
    pc = torch.rand(2, 64, 6)
    tape = model.get_geometry(pc, hp)
    assert [c[0] for c in calls] == ["fps", "gather", "ball"] * 4 + ["three_nn", "weights"] * 4
    assert [c[1] for c in calls if c[0] == "fps"] == [32, 16, 8, 4]
    assert [c for c in calls if c[0] == "three_nn"] == [("three_nn", (2, 8, 3), (2, 4, 3)), ("three_nn", (2, 16, 3), (2, 8, 3)),
                                                         ("three_nn", (2, 32, 3), (2, 16, 3)), ("three_nn", (2, 64, 3), (2, 32, 3))]
    assert [k for k, _, _ in tape.entries] == ["sample"] * 4 + ["interp"] * 4 and len(tape.tensors()) == 16
    n = len(calls)
    xyz = pc[:, :, :3].contiguous()
    with pu.replay_geometry(tape):
        new_xyz, idx = pu.sampling_geometry(32, 0.5, 8, xyz)
        assert new_xyz is tape.entries[0][2][0] and idx is tape.entries[0][2][1] and len(calls) == n  # nothing launched
        with pu.replay_geometry(None):   # e.g. get_geometry of the next batch while a tape is installed
            pu.sampling_geometry(32, 0.5, 8, xyz)
        assert len(calls) == n + 3
        with pytest.raises(RuntimeError, match="out of order"):
            pu.interpolation_geometry(xyz, new_xyz)          # the tape holds layer 2's sampling next
        tape.pos = 1
        with pytest.raises(RuntimeError, match="out of order"):
            pu.sampling_geometry(16, 1.0, 8, xyz)            # right op, wrong input shape
    assert pu._tape is None
    pu.sampling_geometry(32, 0.5, 8, xyz)                    # no tape: computed in place
    assert len(calls) == n + 6
