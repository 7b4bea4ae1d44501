"""CPU: the fp64 layer oracle (oracle/layers_ref.py) against INDEPENDENT implementations.

TensorFlow is absent, so the dense half of the oracle ("parity unpinned" in DESIGN.md section 5)
cannot be checked against the reference itself.  What can be checked is that the restatement computes
what the TF 1.x documentation says those ops compute, using a second, unrelated implementation of the
same mathematics: torch.nn.functional (conv2d / batch_norm / cross_entropy / max_pool), plain numpy,
and finite differences for the backward pass.  These tests pin the oracle the GPU suite relies on.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import layers_ref as lr
from oracle import oracle as orc

D = torch.float64


def _params(rs, scope, k, n):
    p = {}
    lr.init_conv(p, rs, scope, k, n)
    p[scope + "/biases"] = rs.uniform(-0.3, 0.3, n).astype(np.float32)
    p[scope + "/bn/gamma"] = rs.uniform(0.5, 1.5, n).astype(np.float32)
    p[scope + "/bn/beta"] = rs.uniform(-0.3, 0.3, n).astype(np.float32)
    p[scope + "/bn/moving_mean"] = rs.uniform(-0.2, 0.2, n).astype(np.float32)
    p[scope + "/bn/moving_variance"] = rs.uniform(0.5, 2.0, n).astype(np.float32)
    return p


@pytest.mark.parametrize("rank4", [True, False])
def test_conv_bn_relu_train_matches_torch_functional(rank4):
    """tf_util.conv2d/conv1d (1x1, bias, BN, ReLU) in train mode == F.conv + F.batch_norm(eps=1e-3,
    momentum=1-decay) + relu; the moving variance takes the Bessel-corrected batch variance for rank-4
    inputs (TF's fused kernel, which torch.batch_norm shares) and the biased one for rank-3."""
    rs = np.random.RandomState(0)
    k, n, decay = 7, 5, 0.7
    p = _params(rs, "s", k, n)
    x = rs.normal(size=(2, 6, 4, k) if rank4 else (2, 24, k))
    ctx = lr.Ctx(p, is_training=True, bn_decay=decay)
    y = lr.conv_bn_relu(ctx, torch.tensor(x, dtype=D), "s", rank4=rank4)
    w = torch.tensor(p["s/weights"], dtype=D)
    lin = torch.tensor(x, dtype=D) @ w + torch.tensor(p["s/biases"], dtype=D)
    rm = torch.tensor(p["s/bn/moving_mean"], dtype=D).clone()
    rv = torch.tensor(p["s/bn/moving_variance"], dtype=D).clone()
    flat = lin.reshape(-1, n)
    exp = F.batch_norm(flat, rm, rv, torch.tensor(p["s/bn/gamma"], dtype=D),
                       torch.tensor(p["s/bn/beta"], dtype=D), training=True, momentum=1 - decay, eps=1e-3)
    exp = torch.relu(exp).reshape(y.shape)
    np.testing.assert_allclose(y.detach().numpy(), exp.numpy(), atol=1e-12)
    np.testing.assert_allclose(ctx.new_moving["s/bn/moving_mean"], rm.numpy(), atol=1e-12)
    if rank4:  # torch updates running_var with the unbiased variance, like TF's fused batch norm
        np.testing.assert_allclose(ctx.new_moving["s/bn/moving_variance"], rv.numpy(), atol=1e-12)
    else:      # non-fused path: biased variance
        biased = flat.var(0, unbiased=False).numpy()
        np.testing.assert_allclose(ctx.new_moving["s/bn/moving_variance"],
                                   p["s/bn/moving_variance"] * decay + biased * (1 - decay), atol=1e-7)
    # and as a real 1x1 convolution in NCHW, the layout TF would use with use_nchw
    if rank4:
        conv = F.conv2d(torch.tensor(x, dtype=D).permute(0, 3, 1, 2), w.t()[:, :, None, None],
                        torch.tensor(p["s/biases"], dtype=D)).permute(0, 2, 3, 1)
        np.testing.assert_allclose(conv.numpy(), lin.numpy(), atol=1e-12)


def test_conv_bn_eval_mode_uses_moving_statistics():
    rs = np.random.RandomState(1)
    p = _params(rs, "s", 6, 4)
    x = rs.normal(size=(3, 10, 6))
    ctx = lr.Ctx(p, is_training=False)
    y = lr.conv_bn_relu(ctx, torch.tensor(x, dtype=D), "s", rank4=False)
    lin = torch.tensor(x, dtype=D) @ torch.tensor(p["s/weights"], dtype=D) + torch.tensor(p["s/biases"], dtype=D)
    exp = F.batch_norm(lin.reshape(-1, 4), torch.tensor(p["s/bn/moving_mean"], dtype=D),
                       torch.tensor(p["s/bn/moving_variance"], dtype=D), torch.tensor(p["s/bn/gamma"], dtype=D),
                       torch.tensor(p["s/bn/beta"], dtype=D), training=False, eps=1e-3)
    np.testing.assert_allclose(y.detach().numpy(), torch.relu(exp).reshape(y.shape).numpy(), atol=1e-12)
    assert not ctx.new_moving  # eval mode never touches the moving statistics


def test_loss_is_weighted_ce_sum_by_nonzero_weights():
    """tf.losses.sparse_softmax_cross_entropy(labels, logits, weights): sum(w * ce) / #(w != 0)."""
    rs = np.random.RandomState(2)
    pred = rs.normal(size=(2, 50, 9))
    lab = rs.randint(0, 9, (2, 50))
    w = rs.uniform(0.5, 2.0, (2, 50))
    w[0, :11] = 0.0
    got = lr.get_loss(torch.tensor(pred, dtype=D), lab, w)
    ce = F.cross_entropy(torch.tensor(pred, dtype=D).reshape(-1, 9), torch.tensor(lab).reshape(-1),
                         reduction="none")
    exp = (ce * torch.tensor(w, dtype=D).reshape(-1)).sum() / float((w != 0).sum())
    assert abs(got.item() - exp.item()) < 1e-12
    # all-zero weights: the division guard (TF's safe divide) gives 0
    assert lr.get_loss(torch.tensor(pred, dtype=D), lab, np.zeros_like(w)).item() == 0.0


def test_sample_and_group_layout_and_concat_orders():
    """pointnet_util.py:52-54 concatenates [grouped_xyz - centre, features]; the MSG module (:260)
    concatenates [features, xyz]."""
    rs = np.random.RandomState(3)
    xyz = rs.random_sample((2, 200, 3)).astype(np.float32)
    feat = rs.random_sample((2, 200, 4)).astype(np.float32)
    new_xyz, new_points, idx, cnt, fps, gxyz = lr.sample_and_group(16, 0.3, 8, xyz, torch.tensor(feat, dtype=D))
    assert new_points.shape == (2, 16, 8, 7) and idx.shape == (2, 16, 8)
    for b in range(2):
        np.testing.assert_array_equal(new_xyz[b], xyz[b][fps[b]])
        g = xyz[b][idx[b]].astype(np.float64) - new_xyz[b][:, None, :].astype(np.float64)
        np.testing.assert_allclose(new_points[b, :, :, :3].numpy(), g, atol=0)
        np.testing.assert_allclose(new_points[b, :, :, 3:].numpy(), feat[b][idx[b]], atol=0)
    _, msg_order, *_ = lr.sample_and_group(16, 0.3, 8, xyz, torch.tensor(feat, dtype=D), order="feat_first")
    np.testing.assert_allclose(msg_order[..., :4].numpy(), new_points[..., 3:].numpy(), atol=0)
    np.testing.assert_allclose(msg_order[..., 4:].numpy(), new_points[..., :3].numpy(), atol=0)


@pytest.mark.parametrize("pooling", ["max", "avg", "max_and_avg", "weighted_avg"])
def test_sa_module_poolings_against_numpy(pooling):
    """pointnet_util.py:167-191 on top of an identity-free check: recompute the pooled output from the
    oracle's own per-group activations with numpy."""
    rs = np.random.RandomState(4)
    xyz = rs.random_sample((2, 128, 3)).astype(np.float32)
    feat = rs.random_sample((2, 128, 3)).astype(np.float32)
    p = {}
    lr.init_conv(p, rs, "sa/conv0", 6, 8)
    ctx = lr.Ctx(p, is_training=True, bn_decay=0.5)
    new_xyz, out, idx = lr.sa_module(ctx, xyz, torch.tensor(feat, dtype=D), 16, 0.3, 8, [8], "sa", pooling=pooling)
    act = ctx.acts["sa/conv0"].detach().numpy()  # (2,16,8,8) after BN+ReLU
    if pooling == "max":
        exp = act.max(2)
    elif pooling == "avg":
        exp = act.mean(2)
    elif pooling == "max_and_avg":
        exp = np.concatenate([act.mean(2), act.max(2)], -1)  # avg first (pointnet_util.py:185-192)
    else:
        g = xyz[np.arange(2)[:, None, None], idx].astype(np.float64) - new_xyz[:, :, None, :].astype(np.float64)
        e = np.exp(-np.linalg.norm(g, axis=-1, keepdims=True) * 5)
        exp = (act * (e / e.sum(2, keepdims=True))).sum(2)
    np.testing.assert_allclose(out.detach().numpy(), exp, atol=1e-12)


def test_fp_module_weights_and_interpolation():
    """pointnet_util.py:297-309: w_i = (1/max(d_i,1e-10)) / sum_j(1/max(d_j,1e-10)) from the SQUARED
    distances three_nn returns, interpolation = sum_i w_i * points2[idx_i], concat [interp, points1]."""
    rs = np.random.RandomState(5)
    xyz1 = rs.random_sample((2, 60, 3)).astype(np.float32)
    xyz2 = rs.random_sample((2, 12, 3)).astype(np.float32)
    xyz2[0, 0] = xyz1[0, 0]  # an exact coincidence: distance 0 -> floor 1e-10 -> weight ~1
    p1 = rs.random_sample((2, 60, 2))
    p2 = rs.random_sample((2, 12, 5))
    params = {}
    ctx = lr.Ctx(params)
    out = lr.fp_module(ctx, xyz1, xyz2, torch.tensor(p1, dtype=D), torch.tensor(p2, dtype=D), [], "fp")
    dist, idx = orc.three_nn(xyz1, xyz2)
    d = np.maximum(dist.astype(np.float64), 1e-10)
    w = (1.0 / d) / (1.0 / d).sum(-1, keepdims=True)
    np.testing.assert_allclose(lr.fp_weights(dist), w, rtol=3e-7)
    assert lr.fp_weights(dist)[0, 0, 0] > 0.999999
    interp = sum(np.stack([p2[b][idx[b, :, t]] for b in range(2)]) * lr.fp_weights(dist)[..., t:t + 1].astype(np.float64)
                 for t in range(3))
    np.testing.assert_allclose(out[..., :5].numpy(), interp, atol=1e-12)
    np.testing.assert_allclose(out[..., 5:].numpy(), p1, atol=0)


HP_TINY = {"use_color": 1, "l1_npoint": 32, "l1_radius": 0.3, "l1_nsample": 8, "l2_npoint": 16,
           "l2_radius": 0.5, "l2_nsample": 8, "l3_npoint": 8, "l3_radius": 0.8, "l3_nsample": 4,
           "l4_npoint": 4, "l4_radius": 1.2, "l4_nsample": 4}


def test_full_model_shapes_variables_and_finite_difference_gradients():
    """model.py:22-161 end to end on a tiny cloud: every variable of the reference graph exists with the
    reference's scope names, every trainable one receives a gradient, and the autograd gradient of the
    oracle equals a central finite difference of its own loss (so the GPU suite's gradient checks compare
    against a correct derivative)."""
    rs = np.random.RandomState(6)
    pc = np.concatenate([rs.random_sample((2, 96, 3)), rs.random_sample((2, 96, 3))], -1).astype(np.float32)
    lab = rs.randint(0, 9, (2, 96))
    w = rs.uniform(0.5, 1.5, (2, 96))
    params = lr.init_model_params(HP_TINY, 9, seed=1)
    for k in list(params):  # break the symmetry of the default BN/bias initialisation
        if k.endswith("/bn/gamma"):
            params[k] = rs.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
        if k.endswith("/bn/beta") or k.endswith("/biases"):
            params[k] = rs.uniform(-0.3, 0.3, params[k].shape).astype(np.float32)
    names = set(params)
    for l in (1, 2, 3, 4):
        assert "layer%d/conv0/weights" % l in names and "layer%d/conv2/bn/moving_variance" % l in names
    assert {"fa_layer1/conv_0/weights", "fa_layer4/conv_2/bn/gamma", "fc1/weights", "fc1/bn/beta",
            "fc2/weights", "fc2/biases"} <= names and "fc2/bn/gamma" not in names
    assert params["layer1/conv0/weights"].shape == (6, 32) and params["fa_layer4/conv_0/weights"].shape == (131, 128)
    assert params["fa_layer1/conv_0/weights"].shape == (768, 256) and params["fc2/weights"].shape == (128, 9)

    def loss_of(p):
        ctx = lr.Ctx(p, is_training=True, bn_decay=0.5)
        pred = lr.get_model(ctx, pc, 9, HP_TINY)
        assert pred.shape == (2, 96, 9)
        return ctx, lr.get_loss(pred, lab, w)

    ctx, loss = loss_of(params)
    loss.backward()
    grads = ctx.grads()
    trainable = [k for k in params if not k.endswith(("moving_mean", "moving_variance"))]
    assert set(grads) == set(trainable)
    # a conv bias that feeds a train-mode BatchNorm has an exactly zero gradient
    assert np.abs(grads["layer2/conv1/biases"]).max() < 1e-12 and np.abs(grads["fc2/biases"]).max() > 1e-4
    # the loss is only piecewise smooth (ReLU and max-pool switches): with first-layer weights a step of
    # 1e-5 already crosses several kinks (the difference quotient converges to the autograd value only
    # below 1e-7), so the step is 1e-8 -- fp64 leaves ~1e-8 of absolute noise at that size
    eps = 1e-8
    for name, pos in [("fc2/weights", (5, 3)), ("fc1/bn/gamma", (7,)), ("fa_layer4/conv_2/weights", (10, 20)),
                      ("layer1/conv0/weights", (4, 9)), ("layer3/conv1/bn/beta", (30,))]:
        pp = {k: v.astype(np.float64).copy() for k, v in params.items()}
        pp[name][pos] += eps
        lp = loss_of(pp)[1].item()
        pp[name][pos] -= 2 * eps
        lm = loss_of(pp)[1].item()
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - grads[name][pos]) < 3e-6 + 1e-5 * abs(fd), (name, fd, grads[name][pos])


def test_decisions_make_the_oracle_differentiate_the_given_piecewise_function():
    """lr.Ctx(decisions=...): with its own ReLU masks / pooling winners the oracle's outputs and gradients
    are unchanged; a flipped near-zero element changes the gradient path but is accepted as a rounding flip;
    a flip far from the kink is reported by compare_grads."""
    import torch
    from oracle import layers_ref as lr
    rs = np.random.RandomState(5)
    xyz = rs.random_sample((2, 128, 3)).astype(np.float32)
    pts = rs.random_sample((2, 128, 3)).astype(np.float32)
    params = {}
    k = 6
    for i, n in enumerate([8, 16]):
        lr.init_conv(params, rs, "sa/conv%d" % i, k, n)
        k = n

    def run(decisions=None):
        ctx = lr.Ctx(params, is_training=True, bn_decay=0.5, decisions=decisions)
        p = torch.tensor(pts, dtype=torch.float64, requires_grad=True)
        _, out, _ = lr.sa_module(ctx, xyz, p, 16, 0.4, 8, [8, 16], "sa")
        out.sum().backward()
        return ctx, out.detach().numpy(), ctx.grads()

    ctx0, out0, g0 = run()
    own = {}
    for scope in ("sa/conv0", "sa/conv1"):
        own[scope + "/relu_mask"] = (ctx0.acts[scope].detach().numpy() > 0).reshape(-1, ctx0.acts[scope].shape[-1]).astype(np.uint8)
    act = ctx0.acts["sa/conv1"].detach()
    own["sa/conv1/argmax"] = act.argmax(dim=2).numpy().reshape(-1, act.shape[-1]).astype(np.int32)
    ctx1, out1, g1 = run(own)
    np.testing.assert_allclose(out1, out0, atol=1e-12)
    for k2 in g0:
        np.testing.assert_allclose(g1[k2], g0[k2], atol=1e-12)
    assert not lr.compare_grads(ctx1, g0) and all(c == 0 for c, _ in ctx1.flips.values())
    # a decision far from the kink is not a rounding flip
    bad = dict(own)
    m = own["sa/conv0/relu_mask"].copy()
    pre = ctx0.acts["sa/conv0"].detach().numpy().reshape(m.shape)
    r, c = np.unravel_index(np.argmax(pre), pre.shape)
    m[r, c] = 0
    bad["sa/conv0/relu_mask"] = m
    ctx2, _, _ = run(bad)
    msgs = lr.compare_grads(ctx2, g0)
    assert any("not a rounding flip" in s for s in msgs), msgs
