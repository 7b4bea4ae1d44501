"""GPU parity of the PointNet++ layers (pointnet_util / tf_util / model) against the fp64
restatement in oracle/layers_ref.py.  Tolerance: 1e-5 absolute on fp32 forward outputs
(BASELINE.json north_star); gradients are compared with 2e-5 * max(1, |expected|_max)."""
import contextlib

import numpy as np
import pytest

from _util import rng_cloud, to_cuda

pytestmark = pytest.mark.gpu

ATOL = 1e-5


def assert_fwd(got, exp, chained_layers=1):
    """|got - exp| <= 1e-5 * max(1, |exp|) for one module on identical inputs (the north-star bar
    on O(1) values).  For the whole network fed forward through L modules each module amplifies
    the incoming fp32 rounding error (measured in round 1: x1.5-2 per module), so the
    end-to-end bound is 1e-5 * 2^(L-1) capped at 5e-4."""
    exp = np.asarray(exp, np.float64)
    got = np.asarray(got, np.float64)
    bound = min(ATOL * 2.0 ** (chained_layers - 1), 5e-4)
    err = np.abs(got - exp) / np.maximum(1.0, np.abs(exp))
    assert err.max() <= bound, (float(err.max()), bound)


def gtol(exp):
    return 2e-5 * max(1.0, float(np.abs(exp).max()))


@pytest.fixture()
def env(cuda):
    import pn2_b200  # noqa: F401
    from pn2_b200.util import pointnet_util, tf_util
    from pn2_b200 import model
    from oracle import layers_ref as lr
    store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
    return pointnet_util, tf_util, model, lr, store


def load_params(store, params, rank3=("fc1", "fc2")):
    sd = {}
    for k, v in params.items():
        if k.endswith("/weights"):
            scope = k.split("/")[0]
            v = v.reshape((1,) + v.shape) if scope in rank3 else v.reshape((1, 1) + v.shape)
        sd[k] = v
    store.load_state_dict(sd)


def randomize_bn(params, rs):
    for k in list(params):
        if k.endswith("/bn/gamma"):
            params[k] = rs.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
        if k.endswith("/bn/beta") or k.endswith("/biases"):
            params[k] = rs.uniform(-0.3, 0.3, params[k].shape).astype(np.float32)


@contextlib.contextmanager
def recorded_decisions(tf_util):
    """Run the GPU forward inside this block: yields the dict that receives every chain's ReLU masks and
    max-pool winners (tf_util.debug_capture); `as_numpy` turns it into the oracle's `decisions`."""
    tf_util.debug_capture = {}
    try:
        yield tf_util.debug_capture
    finally:
        tf_util.debug_capture = None


def as_numpy(decisions):
    return {k: v.cpu().numpy() for k, v in decisions.items()}


def check_param_grads(store, ctx, names=None, rtol_max=2e-5):
    """Strict elementwise parity (2e-5 * max(1, |g|_max)): the oracle was given the GPU's ReLU masks and
    max-pool winners (each checked to be a legitimate rounding flip), so there is no allowance."""
    from oracle import layers_ref as lr
    assert ctx.decisions, "run the GPU forward under recorded_decisions() and hand them to lr.Ctx"
    ours = {k: v.grad.detach().cpu().numpy() for k, v in store.vars.items() if v.grad is not None}
    bad = lr.compare_grads(ctx, ours, rtol_max=rtol_max)
    if names is not None:
        bad = [b for b in bad if b.split(":")[0] in names]
    assert not bad, "; ".join(bad)


def check_input_grad(got, exp, ctx):
    exp = np.asarray(exp, np.float64)
    got = got.detach().cpu().numpy().astype(np.float64)
    assert ctx.decisions and np.abs(got - exp).max() <= gtol(exp), (np.abs(got - exp).max(), gtol(exp))


def test_sa_module_config1(env):
    """BASELINE.json configs[0]: B=2, N=1024, npoint=256, nsample=32, C=3, radius 0.2,
    mlp [32,32,64], train-mode BN: every intermediate against the oracle."""
    pu, tf_util, _, lr, store = env
    import torch
    rs = np.random.RandomState(100)
    xyz = rs.random_sample((2, 1024, 3)).astype(np.float32)
    pts = rs.random_sample((2, 1024, 3)).astype(np.float32)
    params = {}
    k = 6
    for i, n in enumerate([32, 32, 64]):
        lr.init_conv(params, rs, "layer1/conv%d" % i, k, n)
        k = n
    randomize_bn(params, rs)
    load_params(store, params)
    pt = to_cuda(pts).requires_grad_(True)
    with recorded_decisions(tf_util) as dec:
        new_xyz, out, idx = pu.pointnet_sa_module(to_cuda(xyz), pt, 256, 0.2, 32, [32, 32, 64], None,
                                                  False, True, 0.7, "layer1")
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.7, decisions=as_numpy(dec))
    pts_ref = torch.tensor(pts, dtype=torch.float64, requires_grad=True)
    e_xyz, e_out, e_idx = lr.sa_module(ctx, xyz, pts_ref, 256, 0.2, 32, [32, 32, 64], "layer1")
    np.testing.assert_array_equal(idx.cpu().numpy(), e_idx)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), e_xyz)
    np.testing.assert_allclose(out.detach().cpu().numpy(), e_out.detach().numpy(), atol=ATOL)
    for k2, v in ctx.new_moving.items():
        np.testing.assert_allclose(store.vars[k2].data.cpu().numpy(), v, atol=ATOL, err_msg=k2)

    g = rs.normal(size=tuple(out.shape)).astype(np.float32)
    e_out.backward(torch.tensor(g, dtype=torch.float64))
    out.backward(to_cuda(g))
    check_param_grads(store, ctx)
    check_input_grad(pt.grad, pts_ref.grad.numpy(), ctx)


def test_sample_and_group_intermediates(env):
    pu, _, _, lr, _ = env
    import torch
    rs = np.random.RandomState(3)
    xyz = rs.random_sample((2, 512, 3)).astype(np.float32)
    pts = rs.random_sample((2, 512, 5)).astype(np.float32)
    e_xyz, e_np, e_idx, e_cnt, e_fps, e_gx = lr.sample_and_group(
        64, 0.25, 16, xyz, torch.tensor(pts, dtype=torch.float64))
    new_xyz, new_points, idx, grouped_xyz = pu.sample_and_group(64, 0.25, 16, to_cuda(xyz),
                                                                to_cuda(pts))
    np.testing.assert_array_equal(idx.cpu().numpy(), e_idx)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), e_xyz)
    # exact: the centre subtraction is one fp32 subtraction in both
    exp = e_np.numpy()
    np.testing.assert_allclose(new_points.cpu().numpy(), exp, atol=1e-7)
    np.testing.assert_allclose(grouped_xyz.cpu().numpy(), e_gx.numpy(), atol=1e-7)
    # points=None and use_xyz=False variants
    _, np2, _, _ = pu.sample_and_group(64, 0.25, 16, to_cuda(xyz), None)
    np.testing.assert_allclose(np2.cpu().numpy(), e_gx.numpy(), atol=1e-7)
    _, np3, _, _ = pu.sample_and_group(64, 0.25, 16, to_cuda(xyz), to_cuda(pts), use_xyz=False)
    np.testing.assert_array_equal(np3.cpu().numpy(), exp[..., 3:].astype(np.float32))


def test_sample_and_group_knn_branch(env):
    """knn=True (pointnet_util.py:39-40): the fused kNN kernel feeds the grouping; indices and grouped
    tensors against the oracle, at the reference's own smoke-test shape (test_tf_ops.py:9-24 runs
    knn_point(64, ...) + group_point on (32,512,64) / (32,128,3) without asserting anything)."""
    pu, _, _, lr, _ = env
    import torch
    rs = np.random.RandomState(100)
    xyz = rs.random_sample((4, 512, 3)).astype(np.float32)
    pts = rs.random_sample((4, 512, 64)).astype(np.float32)
    e_xyz, e_np, e_idx, _, _, e_gx = lr.sample_and_group(128, 0.1, 64, xyz, torch.tensor(pts, dtype=torch.float64),
                                                        knn=True)
    new_xyz, new_points, idx, grouped_xyz = pu.sample_and_group(128, 0.1, 64, to_cuda(xyz), to_cuda(pts), knn=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), e_idx)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), e_xyz)
    np.testing.assert_allclose(new_points.cpu().numpy(), e_np.numpy(), atol=1e-7)
    np.testing.assert_allclose(grouped_xyz.cpu().numpy(), e_gx.numpy(), atol=1e-7)
    # every query is a data point: its nearest neighbour is itself
    fps = env[0].farthest_point_sample(128, to_cuda(xyz)).cpu().numpy()
    np.testing.assert_array_equal(idx.cpu().numpy()[:, :, 0], fps)


@pytest.mark.parametrize("pooling,mlp2,group_all", [("avg", None, False), ("max_and_avg", None, False),
                                                    ("weighted_avg", None, False),
                                                    ("max", [48, 24], False), ("max", None, True)])
def test_sa_module_options(env, pooling, mlp2, group_all):
    """avg / weighted_avg / max_and_avg pooling (native kernels), mlp2 and group_all: forward 1e-5 and
    every parameter / input gradient strictly against the fp64 oracle (pointnet_util.py:137-211)."""
    pu, tf_util, _, lr, store = env
    import torch
    rs = np.random.RandomState(17)
    xyz = rs.random_sample((2, 256, 3)).astype(np.float32)
    pts = rs.random_sample((2, 256, 4)).astype(np.float32)
    params = {}
    k = 7
    for i, n in enumerate([16, 32]):
        lr.init_conv(params, rs, "sa/conv%d" % i, k, n)
        k = n
    if pooling == "max_and_avg":
        k *= 2
    for i, n in enumerate(mlp2 or []):
        lr.init_conv(params, rs, "sa/conv_post_%d" % i, k, n)
        k = n
    randomize_bn(params, rs)
    load_params(store, params)
    pt = to_cuda(pts).requires_grad_(True)
    with recorded_decisions(tf_util) as dec:
        _, out, _ = pu.pointnet_sa_module(to_cuda(xyz), pt, 32, 0.3, 16, [16, 32], mlp2,
                                          group_all, True, 0.5, "sa", pooling=pooling)
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.5, decisions=as_numpy(dec))
    pts_ref = torch.tensor(pts, dtype=torch.float64, requires_grad=True)
    _, e_out, _ = lr.sa_module(ctx, xyz, pts_ref, 32, 0.3, 16, [16, 32], "sa", mlp2=mlp2, group_all=group_all,
                               pooling=pooling)
    np.testing.assert_allclose(out.detach().cpu().numpy(), e_out.detach().numpy(), atol=ATOL)
    g = rs.normal(size=tuple(out.shape)).astype(np.float32)
    e_out.backward(torch.tensor(g, dtype=torch.float64))
    out.backward(to_cuda(g))
    check_param_grads(store, ctx)
    check_input_grad(pt.grad, pts_ref.grad.numpy(), ctx)


def test_sa_module_msg(env):
    pu, _, _, lr, store = env
    import torch
    rs = np.random.RandomState(23)
    xyz = rs.random_sample((2, 1024, 3)).astype(np.float32)
    pts = rs.random_sample((2, 1024, 6)).astype(np.float32)
    radii, nss, mlps = [0.1, 0.2, 0.4], [16, 32, 128], [[32, 32, 64], [64, 64, 128], [64, 96, 128]]
    params = {}
    for i, mlp in enumerate(mlps):
        k = 9
        for j, n in enumerate(mlp):
            lr.init_conv(params, rs, "msg/conv%d_%d" % (i, j), k, n)
            k = n
    randomize_bn(params, rs)
    load_params(store, params)
    pt = to_cuda(pts).requires_grad_(True)
    with recorded_decisions(env[1]) as dec:
        new_xyz, out = pu.pointnet_sa_module_msg(to_cuda(xyz), pt, 128, radii, nss, mlps, True, 0.9,
                                                 "msg")
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.9, decisions=as_numpy(dec))
    pts_ref = torch.tensor(pts, dtype=torch.float64, requires_grad=True)
    e_xyz, e_out = lr.sa_module_msg(ctx, xyz, pts_ref, 128, radii, nss, mlps, "msg")
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), e_xyz)
    np.testing.assert_allclose(out.detach().cpu().numpy(), e_out.detach().numpy(), atol=ATOL)
    g = rs.normal(size=tuple(out.shape)).astype(np.float32)
    e_out.backward(torch.tensor(g, dtype=torch.float64))
    out.backward(to_cuda(g))
    check_param_grads(store, ctx)
    check_input_grad(pt.grad, pts_ref.grad.numpy(), ctx)


def test_fp_module(env):
    pu, _, _, lr, store = env
    import torch
    rs = np.random.RandomState(29)
    xyz1 = rs.random_sample((2, 512, 3)).astype(np.float32)
    xyz2 = np.ascontiguousarray(xyz1[:, :64])
    p1 = rs.random_sample((2, 512, 5)).astype(np.float32)
    p2 = rs.random_sample((2, 64, 24)).astype(np.float32)
    params = {}
    k = 29
    for i, n in enumerate([64, 32]):
        lr.init_conv(params, rs, "fa/conv_%d" % i, k, n)
        k = n
    randomize_bn(params, rs)
    load_params(store, params)
    t1, t2 = to_cuda(p1).requires_grad_(True), to_cuda(p2).requires_grad_(True)
    with recorded_decisions(env[1]) as dec:
        out = pu.pointnet_fp_module(to_cuda(xyz1), to_cuda(xyz2), t1, t2, [64, 32], True, 0.9, "fa")
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.9, decisions=as_numpy(dec))
    r1 = torch.tensor(p1, dtype=torch.float64, requires_grad=True)
    r2 = torch.tensor(p2, dtype=torch.float64, requires_grad=True)
    e_out = lr.fp_module(ctx, xyz1, xyz2, r1, r2, [64, 32], "fa")
    np.testing.assert_allclose(out.detach().cpu().numpy(), e_out.detach().numpy(), atol=ATOL)
    g = rs.normal(size=tuple(out.shape)).astype(np.float32)
    e_out.backward(torch.tensor(g, dtype=torch.float64))
    out.backward(to_cuda(g))
    check_param_grads(store, ctx)
    check_input_grad(t1.grad, r1.grad.numpy(), ctx)
    check_input_grad(t2.grad, r2.grad.numpy(), ctx)
    # points1=None branch
    store2 = env[1].set_default_store(env[1].VariableStore(device="cuda"))
    params2 = {}
    k = 24
    for i, n in enumerate([16]):
        lr.init_conv(params2, rs, "fb/conv_%d" % i, k, n)
    load_params(store2, params2)
    ctx2 = lr.Ctx(params2, is_training=True, bn_decay=0.9)
    e2 = lr.fp_module(ctx2, xyz1, xyz2, None, torch.tensor(p2, dtype=torch.float64), [16], "fb")
    o2 = pu.pointnet_fp_module(to_cuda(xyz1), to_cuda(xyz2), None, to_cuda(p2), [16], True, 0.9,
                               "fb")
    np.testing.assert_allclose(o2.detach().cpu().numpy(), e2.detach().numpy(), atol=ATOL)


HP_SMALL = {"use_color": 1, "l1_npoint": 256, "l1_radius": 0.1, "l1_nsample": 32,
            "l2_npoint": 64, "l2_radius": 0.2, "l2_nsample": 32, "l3_npoint": 16,
            "l3_radius": 0.4, "l3_nsample": 32, "l4_npoint": 8, "l4_radius": 0.8,
            "l4_nsample": 32}


def run_model_parity(env, hp, b, n, scale, train=True):
    pu, tf_util, model, lr, store = env
    import torch
    rs = np.random.RandomState(100)
    pc = np.concatenate([rs.random_sample((b, n, 3)) * np.asarray(scale),
                         rs.random_sample((b, n, 3))], -1).astype(np.float32)
    labels = rs.randint(0, 9, (b, n)).astype(np.int32)
    smpw = rs.uniform(0.5, 2.0, (b, n)).astype(np.float32)
    smpw[0, :7] = 0.0
    params = lr.init_model_params(hp, 9, seed=1)
    randomize_bn(params, rs)
    load_params(store, params)
    seed = 1234
    tf_util.set_dropout_seed(seed)
    mask = tf_util.dropout_mask(b * n * 128, 0.5, seed).cpu().numpy().reshape(b, n, 128)
    with recorded_decisions(tf_util) as dec:
        pred, end_points = model.get_model(to_cuda(pc), train, 9, hp, bn_decay=0.5)
    ctx = lr.Ctx(params, is_training=train, bn_decay=0.5,
                 dropout_masks={"dp1": mask.astype(np.float64)}, decisions=as_numpy(dec))
    e_pred = lr.get_model(ctx, pc, 9, hp)
    assert_fwd(pred.detach().cpu().numpy(), e_pred.detach().numpy(), chained_layers=1 if not train else 10)
    if not train:
        return
    e_loss = lr.get_loss(e_pred, labels, smpw)
    loss = model.get_loss(pred, to_cuda(labels), to_cuda(smpw), end_points)
    assert abs(loss.item() - e_loss.item()) < 5e-5
    e_loss.backward()
    loss.backward()
    # whole network: the gradient of the first SA layer has crossed ~50 fp32 layer passes (26 forward, 26
    # backward); measured worst case 2.4e-5 * |g|_max (B200, strict comparison, no flip allowance), single
    # modules stay below 2e-5 -- the bound for the chained network is 5e-5
    check_param_grads(store, ctx, rtol_max=5e-5)
    for k2, v in ctx.new_moving.items():
        np.testing.assert_allclose(store.vars[k2].data.cpu().numpy(), v, atol=1e-4, rtol=1e-4, err_msg=k2)


def test_full_model_small(env):
    run_model_parity(env, HP_SMALL, 2, 1024, (1.0, 1.0, 1.0))


def test_full_model_eval_mode(env):
    run_model_parity(env, HP_SMALL, 2, 1024, (1.0, 1.0, 1.0), train=False)


def test_full_model_semantic_json_shape(env):
    """semantic.json hyper-parameters (config 2 of BASELINE.json) at B=2: N=8192 points in a
    10 x 10 x 5 box, radii 0.5/1/2/4, npoint 1024/256/64/16."""
    hp = {"use_color": 1, "l1_npoint": 1024, "l1_radius": 0.5, "l1_nsample": 32,
          "l2_npoint": 256, "l2_radius": 1.0, "l2_nsample": 32, "l3_npoint": 64,
          "l3_radius": 2.0, "l3_nsample": 32, "l4_npoint": 16, "l4_radius": 4.0,
          "l4_nsample": 32}
    run_model_parity(env, hp, 2, 8192, (10.0, 10.0, 5.0))


def test_variable_names_match_reference(env):
    """Scopes of model.py:36-146 + tf_util.py:98-111,172-199 give the checkpoint names."""
    _, _, model, _, store = env
    import torch
    pc = to_cuda(np.random.RandomState(0).random_sample((1, 512, 6)).astype(np.float32))
    model.get_model(pc, True, 9, HP_SMALL, bn_decay=0.9)
    names = set(store.vars)
    for must in ["layer1/conv0/weights", "layer1/conv0/biases", "layer1/conv0/bn/gamma",
                 "layer1/conv0/bn/beta", "layer1/conv0/bn/moving_mean",
                 "layer1/conv0/bn/moving_variance", "layer4/conv2/weights",
                 "fa_layer1/conv_0/weights", "fa_layer4/conv_2/bn/gamma", "fc1/weights",
                 "fc1/bn/beta", "fc2/weights", "fc2/biases"]:
        assert must in names, must
    assert "fc2/bn/gamma" not in names
    assert tuple(store.vars["layer1/conv0/weights"].data.shape) == (1, 1, 6, 32)
    assert tuple(store.vars["fc2/weights"].data.shape) == (1, 128, 9)
    n_train = sum(v.data.numel() for v in store.trainable())
    # the all-reduce message: 967 945 fp32 with the reference's 3 colour channels (SURVEY.md 3.1 quotes
    # 968 425, which is the same network with BASELINE.json's 6 feature channels: +96 +384 weights)
    assert n_train == 967945


def test_predictor_on_gpu(env, tmp_path):
    """predict.Predictor (predict.py:15-105) end to end on the GPU: checkpoint dict / .npz (with the
    optimizer slots a TF Saver also writes) -> eval-mode forward (moving-average BatchNorm, no dropout) ->
    arg-max labels; then the dense label transfer.  Labels against the fp64 oracle wherever the oracle's
    top-2 logit margin exceeds the forward tolerance; a checkpoint lacking a variable must fail loudly."""
    _, tf_util, _, lr, _ = env
    import pn2_b200
    from pn2_b200 import predict
    from oracle import oracle as orc
    rs = np.random.RandomState(7)
    hp = dict(HP_SMALL)
    params = lr.init_model_params(hp, 9, seed=3)
    randomize_bn(params, rs)
    for k in list(params):
        if k.endswith("moving_mean"):
            params[k] = rs.normal(0, 0.2, params[k].shape).astype(np.float32)
        if k.endswith("moving_variance"):
            params[k] = rs.uniform(0.5, 2.0, params[k].shape).astype(np.float32)
    ckpt = dict(params)
    ckpt["layer1/conv0/weights/Adam"] = np.zeros_like(params["layer1/conv0/weights"])
    ckpt["beta1_power"] = np.float32(0.9)
    path = str(tmp_path / "ckpt.npz")
    np.savez(path, **ckpt)
    p = predict.Predictor(path, 9, hp)
    assert sorted(p.skipped) == ["beta1_power", "layer1/conv0/weights/Adam"]
    pc = np.concatenate([rs.random_sample((2, 1024, 3)), rs.random_sample((2, 1024, 3))], -1).astype(np.float32)
    labels = p.predict(pc)
    ctx = lr.Ctx(params, is_training=False)
    e_pred = lr.get_model(ctx, pc, 9, hp).detach().numpy()
    top2 = np.sort(e_pred, -1)[..., -2:]
    sure = (top2[..., 1] - top2[..., 0]) > 1e-3
    assert sure.mean() > 0.9
    np.testing.assert_array_equal(labels[sure], e_pred.argmax(-1)[sure])
    assert labels.shape == (2, 1024) and labels.dtype == np.int64
    # dense label transfer (predict.py:93-105): sparse cloud = the first cloud, labels = the prediction
    dense = rs.random_sample((5000, 3)).astype(np.float32)
    dl, dc = p.interpolate_labels(pc[0, :, :3], labels[0], dense, knn=3)
    el, ec = orc.interpolate_label_with_color(pc[0, :, :3], labels[0].astype(np.int32), dense, 3)
    np.testing.assert_array_equal(dl, el)
    np.testing.assert_array_equal(dc, ec)
    # a variable missing from the checkpoint: KeyError at predict time, not a silent xavier init
    broken = {k: v for k, v in params.items() if not k.startswith("fa_layer3/conv_1/")}
    with pytest.raises(KeyError, match="not in the loaded checkpoint"):
        predict.Predictor(broken, 9, hp).predict(pc)
    with pytest.raises(ValueError, match="batch_data must be"):
        p.predict(pc[:, :, :3])


def test_full_model_every_module_meets_1e5_on_identical_inputs(env, monkeypatch):
    """The north-star bar (1e-5 abs) module by module INSIDE the full network: every SA / FP module and the head
    of one training forward is re-evaluated by the fp64 oracle on the GPU's own inputs to that module, so the
    fp32 rounding of earlier modules (which the chained comparison has to allow for) drops out.  semantic.json
    radii / npoint / nsample at B=2 x 2048 points."""
    pu, tf_util, model, lr, store = env
    import torch
    hp = {"use_color": 1, "l1_npoint": 512, "l1_radius": 0.5, "l1_nsample": 32, "l2_npoint": 128, "l2_radius": 1.0,
          "l2_nsample": 32, "l3_npoint": 32, "l3_radius": 2.0, "l3_nsample": 32, "l4_npoint": 8, "l4_radius": 4.0,
          "l4_nsample": 32}
    rs = np.random.RandomState(100)
    b, n = 2, 2048
    pc = np.concatenate([rs.random_sample((b, n, 3)) * [10.0, 10.0, 5.0], rs.random_sample((b, n, 3))], -1).astype(np.float32)
    params = lr.init_model_params(hp, 9, seed=1)
    randomize_bn(params, rs)
    load_params(store, params)
    calls = []
    real_sa, real_fp, real_conv1d = model.pointnet_sa_module, model.pointnet_fp_module, tf_util.conv1d

    def sa(xyz, points, **kw):
        out = real_sa(xyz, points, **kw)
        calls.append(("sa", kw["scope"], (xyz, points), kw, out))
        return out

    def fp(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope):
        out = real_fp(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope=scope)
        calls.append(("fp", scope, (xyz1, xyz2, points1, points2), {"mlp": mlp}, out))
        return out

    def conv1d(inputs, num_output_channels, kernel_size, scope, **kw):
        out = real_conv1d(inputs, num_output_channels, kernel_size, scope=scope, **kw)
        calls.append(("conv1d", scope, (inputs,), kw, out))
        return out

    monkeypatch.setattr(model, "pointnet_sa_module", sa)
    monkeypatch.setattr(model, "pointnet_fp_module", fp)
    monkeypatch.setattr(tf_util, "conv1d", conv1d)
    with torch.no_grad():
        model.get_model(to_cuda(pc), True, 9, hp, bn_decay=0.5)
    assert [c[1] for c in calls] == ["layer1", "layer2", "layer3", "layer4", "fa_layer1", "fa_layer2", "fa_layer3",
                                     "fa_layer4", "fc1", "fc2"]
    npy = lambda t: None if t is None else t.detach().cpu().numpy()  # noqa: E731
    t64 = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64)  # noqa: E731
    worst = {}
    for kind, scope, ins, kw, out in calls:
        ctx = lr.Ctx(params, is_training=True, bn_decay=0.5)
        if kind == "sa":
            xyz, points = npy(ins[0]), npy(ins[1])
            e_xyz, e_out, e_idx = lr.sa_module(ctx, xyz, t64(points), kw["npoint"], kw["radius"], kw["nsample"],
                                               kw["mlp"], scope)
            np.testing.assert_array_equal(npy(out[2]), e_idx, err_msg=scope)
            np.testing.assert_array_equal(npy(out[0]), e_xyz, err_msg=scope)
            got, exp = npy(out[1]), e_out.detach().numpy()
        elif kind == "fp":
            e = lr.fp_module(ctx, npy(ins[0]), npy(ins[1]), t64(npy(ins[2])), t64(npy(ins[3])), kw["mlp"], scope)
            got, exp = npy(out), e.detach().numpy()
        else:
            x = t64(npy(ins[0]))
            e = lr.conv_bn_relu(ctx, x, scope, bn=kw.get("bn", False), relu=kw.get("activation_fn", "relu") is not None,
                                rank4=False)
            got, exp = npy(out), e.detach().numpy()
        err = float((np.abs(got - exp) / np.maximum(1.0, np.abs(exp))).max())
        worst[scope] = err
        assert err <= ATOL, (scope, err)
    print("per-module max error on identical inputs:", {k: "%.2g" % v for k, v in worst.items()})
