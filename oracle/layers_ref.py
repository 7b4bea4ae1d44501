"""fp64 restatement of the reference's PointNet++ layers (the floating-point part).

TEST INFRASTRUCTURE ONLY (see oracle/pn2_oracle.c header).  Index-valued ops
(FPS, ball query, 3-NN) come from the C oracle; the dense math is restated with
torch CPU float64 tensors so that ``torch.autograd`` supplies the backward pass
the reference gets from TF autodiff.

Follows (reference file:line):
  sample_and_group          util/pointnet_util.py:18-60   (concat order [xyz, feat], :52-54)
  sample_and_group_all      util/pointnet_util.py:63-95
  pointnet_sa_module        util/pointnet_util.py:98-216
  pointnet_sa_module_msg    util/pointnet_util.py:219-282 (concat order [feat, xyz], :260)
  pointnet_fp_module        util/pointnet_util.py:285-326 (weights from squared dist, :300-303)
  conv2d / conv1d (1x1)     util/tf_util.py:128-204, 54-125  (bias, then BN, then ReLU)
  batch_norm_template       util/tf_util.py:555-581  -> tf.contrib.layers.batch_norm
  dropout                   util/tf_util.py:646-665
  get_model / get_loss      model.py:22-161

PARITY UNPINNED: no reference test pins these results and TensorFlow is not
vendored.  TF-side facts restated from TF 1.x semantics: batch_norm epsilon 1e-3,
biased batch variance for normalisation, moving statistics updated as
``moving -= (moving - batch) * (1 - decay)``; for rank-4 inputs (conv2d) the
fused kernel feeds the Bessel-corrected variance to the moving average, rank-3
(conv1d) uses the biased one; tf.losses.sparse_softmax_cross_entropy reduces with
SUM_BY_NONZERO_WEIGHTS; tf.nn.dropout scales kept values by 1/keep_prob.
"""
import numpy as np
import torch

from . import oracle as orc

BN_EPS = 1e-3
F64 = torch.float64  # working dtype; bench.py's CPU-baseline leg switches it to float32


def set_dtype(dt):
    """fp64 for parity checks (default); fp32 when the module is TIMED as the CPU baseline."""
    global F64
    F64 = dt


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=F64)


def xavier_uniform(rng, k, n):
    """tf.contrib.layers.xavier_initializer() for a [1,1,k,n] kernel (tf_util.py:43-47)."""
    lim = np.sqrt(6.0 / (k + n))
    return rng.uniform(-lim, lim, size=(k, n)).astype(np.float32)


def init_conv(params, rng, scope, k, n, bn=True):
    """Variables tf_util.conv2d/conv1d would create under ``scope``."""
    params[scope + "/weights"] = xavier_uniform(rng, k, n)
    params[scope + "/biases"] = np.zeros(n, np.float32)
    if bn:
        params[scope + "/bn/gamma"] = np.ones(n, np.float32)
        params[scope + "/bn/beta"] = np.zeros(n, np.float32)
        params[scope + "/bn/moving_mean"] = np.zeros(n, np.float32)
        params[scope + "/bn/moving_variance"] = np.ones(n, np.float32)


class Ctx:
    """Holds fp64 leaf tensors for every trainable variable + BN moving stats."""

    def __init__(self, params, is_training=True, bn_decay=None, dropout_masks=None, decisions=None):
        self.np_params = params
        self.t = {}
        for k, v in params.items():
            trainable = not (k.endswith("moving_mean") or k.endswith("moving_variance"))
            self.t[k] = _t(v).clone().requires_grad_(trainable)
        self.is_training = is_training
        self.bn_decay = 0.9 if bn_decay is None else bn_decay
        self.new_moving = {}
        self.dropout_masks = dropout_masks or {}
        self.acts = {}
        self.near_zero = {}
        # decisions of the implementation under test at the non-differentiable points of the network:
        # "<scope>/relu_mask" (M,N) 0/1 and "<scope>/argmax" (G,N) winners of the max-pool.  When given,
        # the oracle evaluates ReLU and max-pool WITH those decisions (after checking that each one is
        # legitimate: a flipped element lies within 1e-4 of the kink, a different winner within 1e-4 of
        # the maximum), so both sides differentiate the same piecewise-linear function and gradients can
        # be compared elementwise with no allowance for flips.
        self.decisions = {k: np.asarray(v) for k, v in (decisions or {}).items()}
        self.flips = {}       # scope -> (count of decisions that differ from fp64's, largest |value| among them)

    def grads(self):
        return {k: v.grad.double().numpy().copy() for k, v in self.t.items() if v.grad is not None}


def conv_bn_relu(ctx, x, scope, bn=True, relu=True, rank4=True):
    """1x1 conv == x @ W + b over the last axis, then BN, then ReLU."""
    w, b = ctx.t[scope + "/weights"], ctx.t[scope + "/biases"]
    y = x @ w + b
    if bn:
        g, beta = ctx.t[scope + "/bn/gamma"], ctx.t[scope + "/bn/beta"]
        mm, mv = ctx.t[scope + "/bn/moving_mean"], ctx.t[scope + "/bn/moving_variance"]
        if ctx.is_training:
            flat = y.reshape(-1, y.shape[-1])
            mean = flat.mean(0)
            var = flat.var(0, unbiased=False)
            cnt = flat.shape[0]
            upd_var = var * (cnt / max(cnt - 1, 1)) if rank4 else var
            d = ctx.bn_decay
            ctx.new_moving[scope + "/bn/moving_mean"] = (
                mm - (mm - mean.detach()) * (1 - d)).detach().numpy().copy()
            ctx.new_moving[scope + "/bn/moving_variance"] = (
                mv - (mv - upd_var.detach()) * (1 - d)).detach().numpy().copy()
        else:
            mean, var = mm, mv
        y = (y - mean) * torch.rsqrt(var + BN_EPS) * g + beta
    if relu:
        # ReLU-boundary census: an element whose pre-activation is within fp32 noise of zero can
        # be masked differently by an fp32 implementation; one such flip shifts the BatchNorm
        # gradient sums of its channel by O(upstream gradient) (see compare_grads)
        ctx.near_zero[scope] = int((y.detach().abs() < 1e-5).sum())
        m = ctx.decisions.get(scope + "/relu_mask")
        if m is not None:
            mt = torch.as_tensor(m.astype(bool)).reshape(y.shape)
            diff = mt != (y.detach() > 0)
            ctx.flips[scope] = (int(diff.sum()), float(y.detach().abs()[diff].max()) if bool(diff.any()) else 0.0)
            y = y * mt.to(y.dtype)
        else:
            y = torch.relu(y)
    ctx.acts[scope] = y
    return y


def group_rows(x, idx):
    """group_point: x (B,N,C), idx (B,m,ns) -> (B,m,ns,C); differentiable gather."""
    b, m, ns = idx.shape
    ii = torch.as_tensor(np.asarray(idx), dtype=torch.int64).reshape(b, m * ns)
    out = torch.gather(x, 1, ii[:, :, None].expand(b, m * ns, x.shape[2]))
    return out.reshape(b, m, ns, x.shape[2])


def sample_and_group(npoint, radius, nsample, xyz_np, points, use_xyz=True, order="xyz_first", knn=False):
    fps = orc.farthest_point_sample(npoint, xyz_np)
    new_xyz_np = orc.gather_point(xyz_np, fps)
    if knn:  # pointnet_util.py:39-40
        _, idx = orc.knn_point(nsample, xyz_np, new_xyz_np)
        cnt = None
    else:
        idx, cnt = orc.query_ball_point(radius, nsample, xyz_np, new_xyz_np)
    xyz = _t(xyz_np)
    grouped_xyz = group_rows(xyz, idx) - _t(new_xyz_np)[:, :, None, :]
    if points is not None:
        gp = group_rows(points, idx)
        if use_xyz:
            new_points = (torch.cat([grouped_xyz, gp], -1) if order == "xyz_first"
                          else torch.cat([gp, grouped_xyz], -1))
        else:
            new_points = gp
    else:
        new_points = grouped_xyz
    return new_xyz_np, new_points, idx, cnt, fps, grouped_xyz


def _max_pool(ctx, x, scope):
    """max over dim 2 of (B,m,ns,C); with a recorded argmax the winners of the implementation under test
    are used (and checked to be within 1e-4 of the true maximum)."""
    arg = ctx.decisions.get(scope + "/argmax")
    best = x.max(dim=2, keepdim=True)[0]
    if arg is None:
        return best
    b, m, ns, c = x.shape
    a = torch.as_tensor(arg.astype(np.int64)).reshape(b, m, 1, c)
    picked = torch.gather(x, 2, a)
    gap = (best - picked).detach()
    ctx.flips[scope + "/argmax"] = (int((gap > 0).sum()), float(gap.max()))
    return picked


def sa_module(ctx, xyz_np, points, npoint, radius, nsample, mlp, scope, mlp2=None,
              group_all=False, bn=True, pooling="max", use_xyz=True):
    if group_all:
        b, n, _ = xyz_np.shape
        new_xyz_np = np.zeros((b, 1, 3), np.float32)
        idx = np.tile(np.arange(n, dtype=np.int32).reshape(1, 1, n), (b, 1, 1))
        grouped_xyz = _t(xyz_np).reshape(b, 1, n, 3)
        if points is not None:
            new_points = (torch.cat([_t(xyz_np), points], 2) if use_xyz else points)[:, None]
        else:
            new_points = grouped_xyz
    else:
        new_xyz_np, new_points, idx, _, _, grouped_xyz = sample_and_group(
            npoint, radius, nsample, xyz_np, points, use_xyz)
    for i, _ in enumerate(mlp):
        new_points = conv_bn_relu(ctx, new_points, "%s/conv%d" % (scope, i), bn=bn)
    if pooling == "max":
        new_points = _max_pool(ctx, new_points, "%s/conv%d" % (scope, len(mlp) - 1))
    elif pooling == "avg":
        new_points = new_points.mean(dim=2, keepdim=True)
    elif pooling == "weighted_avg":
        dists = torch.linalg.norm(grouped_xyz, dim=-1, keepdim=True)
        e = torch.exp(-dists * 5)
        new_points = (new_points * (e / e.sum(dim=2, keepdim=True))).sum(dim=2, keepdim=True)
    elif pooling == "max_and_avg":
        new_points = torch.cat([new_points.mean(dim=2, keepdim=True),
                                _max_pool(ctx, new_points, "%s/conv%d" % (scope, len(mlp) - 1))], -1)
    if mlp2 is not None:
        for i, _ in enumerate(mlp2):
            new_points = conv_bn_relu(ctx, new_points, "%s/conv_post_%d" % (scope, i), bn=bn)
    return new_xyz_np, new_points.squeeze(2), idx


def sa_module_msg(ctx, xyz_np, points, npoint, radius_list, nsample_list, mlp_list, scope,
                  bn=True, use_xyz=True):
    fps = orc.farthest_point_sample(npoint, xyz_np)
    new_xyz_np = orc.gather_point(xyz_np, fps)
    xyz = _t(xyz_np)
    outs = []
    for i, radius in enumerate(radius_list):
        idx, _ = orc.query_ball_point(radius, nsample_list[i], xyz_np, new_xyz_np)
        gx = group_rows(xyz, idx) - _t(new_xyz_np)[:, :, None, :]
        if points is not None:
            g = group_rows(points, idx)
            if use_xyz:
                g = torch.cat([g, gx], -1)
        else:
            g = gx
        for j, _ in enumerate(mlp_list[i]):
            g = conv_bn_relu(ctx, g, "%s/conv%d_%d" % (scope, i, j), bn=bn)
        outs.append(_max_pool(ctx, g, "%s/conv%d_%d" % (scope, i, len(mlp_list[i]) - 1)).squeeze(2))
    return new_xyz_np, torch.cat(outs, -1)


def fp_weights(dist):
    """pointnet_util.py:300-303 in fp32 (the reference computes them in fp32)."""
    d = np.maximum(np.asarray(dist, np.float32), np.float32(1e-10))
    inv = (np.float32(1.0) / d).astype(np.float32)
    norm = ((inv[..., 0] + inv[..., 1]) + inv[..., 2]).astype(np.float32)[..., None]
    return (inv / norm).astype(np.float32)


def fp_module(ctx, xyz1_np, xyz2_np, points1, points2, mlp, scope, bn=True):
    dist, idx = orc.three_nn(xyz1_np, xyz2_np)
    w = _t(fp_weights(dist))
    b, n, _ = idx.shape
    ii = torch.as_tensor(idx, dtype=torch.int64)
    rows = torch.gather(points2, 1, ii.reshape(b, n * 3)[:, :, None].expand(b, n * 3,
                                                                             points2.shape[2]))
    rows = rows.reshape(b, n, 3, -1)
    interp = (rows[:, :, 0] * w[:, :, 0:1] + rows[:, :, 1] * w[:, :, 1:2]) + rows[:, :, 2] * w[:, :, 2:3]
    new_points1 = torch.cat([interp, points1], 2) if points1 is not None else interp
    for i, _ in enumerate(mlp):
        new_points1 = conv_bn_relu(ctx, new_points1, "%s/conv_%d" % (scope, i), bn=bn)
    return new_points1


SA_MLPS = {1: [32, 32, 64], 2: [64, 64, 128], 3: [128, 128, 256], 4: [256, 256, 512]}
FP_MLPS = {1: [256, 256], 2: [256, 256], 3: [256, 128], 4: [128, 128, 128]}


def init_model_params(hp, num_class, seed=0):
    """Create every variable model.get_model (model.py:22-148) would create."""
    rng = np.random.RandomState(seed)
    p = {}
    feat = 3 * int(hp["use_color"]) if hp["use_color"] else 0
    c_in = feat
    sa_out = {0: feat}
    for l in (1, 2, 3, 4):
        k = c_in + 3
        for i, n in enumerate(SA_MLPS[l]):
            init_conv(p, rng, "layer%d/conv%d" % (l, i), k, n)
            k = n
        c_in = k
        sa_out[l] = k
    # fa_layer1: points1 = l3 (256), points2 = l4 (512) ...
    up = sa_out[4]
    skips = {1: sa_out[3], 2: sa_out[2], 3: sa_out[1], 4: feat}
    for l in (1, 2, 3, 4):
        k = up + skips[l]
        for i, n in enumerate(FP_MLPS[l]):
            init_conv(p, rng, "fa_layer%d/conv_%d" % (l, i), k, n)
            k = n
        up = k
    init_conv(p, rng, "fc1", up, 128)
    init_conv(p, rng, "fc2", 128, num_class, bn=False)
    return p


def get_model(ctx, point_cloud_np, num_class, hp):
    """model.py:22-148.  Returns logits (B,N,num_class) fp64."""
    pc = np.asarray(point_cloud_np, np.float32)
    if hp["use_color"]:
        feat = 3 * int(hp["use_color"])
        l0_xyz = np.ascontiguousarray(pc[:, :, :3])
        l0_points = _t(pc[:, :, 3:3 + feat])
    else:
        l0_xyz, l0_points = pc, None
    xyz = {0: l0_xyz}
    pts = {0: l0_points}
    for l in (1, 2, 3, 4):
        xyz[l], pts[l], _ = sa_module(ctx, xyz[l - 1], pts[l - 1], hp["l%d_npoint" % l],
                                      hp["l%d_radius" % l], hp["l%d_nsample" % l], SA_MLPS[l],
                                      "layer%d" % l)
    up = pts[4]
    for l, (lo, hi) in zip((1, 2, 3, 4), ((3, 4), (2, 3), (1, 2), (0, 1))):
        up = fp_module(ctx, xyz[lo], xyz[hi], pts[lo], up, FP_MLPS[l], "fa_layer%d" % l)
    net = conv_bn_relu(ctx, up, "fc1", bn=True, rank4=False)
    ctx.acts["feats"] = net
    if ctx.is_training:
        mask = ctx.dropout_masks.get("dp1")
        if mask is not None:
            net = net * _t(mask) / 0.5
    net = conv_bn_relu(ctx, net, "fc2", bn=False, relu=False)
    return net


def get_loss(pred, label, smpw):
    """model.py:152-161: weighted sparse softmax CE, SUM_BY_NONZERO_WEIGHTS."""
    logp = torch.log_softmax(pred, -1)
    lab = torch.as_tensor(np.asarray(label), dtype=torch.int64)
    ce = -torch.gather(logp, 2, lab[:, :, None]).squeeze(2)
    w = _t(smpw)
    nz = (w != 0).sum().clamp(min=1).to(F64)
    return (ce * w).sum() / nz


def compare_grads(ctx, ours, rtol_max=2e-5, flip_rel_l2=5e-2):
    """Gradient parity check.  ``ours``: name -> numpy array.

    With ``ctx.decisions`` (the ReLU masks / max-pool winners of the implementation under test) the
    oracle differentiated the same piecewise-linear function, so every gradient must agree within
    ``rtol_max * max(1, |g|_max)`` elementwise -- no allowance; in addition every decision that differs
    from fp64's own must be legitimate (the element within 1e-4 of the kink / of the maximum).

    Without decisions (legacy callers) a mismatch is tolerated only if the oracle saw pre-activations
    within 1e-5 of the ReLU kink and the relative L2 error stays below ``flip_rel_l2``.
    Returns a list of failure strings (empty == pass)."""
    bad = []
    strict = bool(ctx.decisions)
    for scope, (cnt, mag) in ctx.flips.items():
        if mag > 1e-4:
            bad.append("%s: %d decisions differ from fp64, one by %.3g (> 1e-4: not a rounding flip)"
                       % (scope, cnt, mag))
    boundary = sum(ctx.near_zero.values())
    for name, e in ctx.grads().items():
        if name not in ours:
            bad.append("%s: missing" % name)
            continue
        got = np.asarray(ours[name], np.float64).reshape(e.shape)
        d = float(np.abs(got - e).max())
        tol = rtol_max * max(1.0, float(np.abs(e).max()))
        if d <= tol:
            continue
        rel = float(np.linalg.norm(got - e) / max(np.linalg.norm(e), 1e-30))
        if not strict and boundary > 0 and rel < flip_rel_l2:
            continue
        bad.append("%s: |diff| %.3g > tol %.3g, rel-L2 %.3g, boundary elements %d%s"
                   % (name, d, tol, rel, boundary, " (strict: decisions fed)" if strict else ""))
    return bad
