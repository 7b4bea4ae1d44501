"""PointNet++ layers: the reference's util/pointnet_util.py signatures on the sm_100a engine.

  sample_and_group(npoint, radius, nsample, xyz, points, knn, use_xyz)   pointnet_util.py:18-60
  sample_and_group_all(xyz, points, use_xyz)                             pointnet_util.py:63-95
  pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2,
                     group_all, is_training, bn_decay, scope, bn, pooling,
                     knn, use_xyz, use_nchw)                             pointnet_util.py:98-216
  pointnet_sa_module_msg(...)                                            pointnet_util.py:219-282
  pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training,
                     bn_decay, scope, bn)                                pointnet_util.py:285-326

Tensors are CUDA float32, channels-last, exactly the reference's layouts.  The common
configuration (ball query, max pooling, no mlp2) runs as: FPS -> gather -> ball query ->
fused group+centre+concat -> shared-MLP chain with BN statistics in the GEMM epilogue ->
fused BN+ReLU+max-pool.  Rarely used options (avg / weighted_avg / max_and_avg pooling,
mlp2, group_all) are composed from the same kernels plus elementwise torch glue.
"""
import torch

from .._ffi import F32, I32, call, ptr
from ..tf_ops.tf_grouping import group_point, knn_point, query_ball_point
from ..tf_ops.tf_interpolate import three_nn
from ..tf_ops.tf_sampling import farthest_point_sample, gather_point
from . import tf_util


def _pad4(w):
    return (w + 3) // 4 * 4


# ---- geometry tape ---------------------------------------------------------------------------------
# Farthest point sampling, gather_point, ball query, three_nn and the interpolation weights depend on the
# point coordinates only, never on the weights.  `sampling_geometry` / `interpolation_geometry` are the two
# places the layers compute them; with a GeometryTape set (`replay_geometry`) the layers take the precomputed
# tensors from it instead, in call order -- which lets train_step.Trainer run the geometry of the NEXT batch
# on a second stream next to the dense stage of the current one (model.get_geometry builds the tape).
class GeometryTape:
    """Results of the weight-independent ops of one forward pass, in the order the layers ask for them."""

    def __init__(self):
        self.entries, self.pos = [], 0

    def add(self, kind, key, value):
        self.entries.append((kind, tuple(key), tuple(value)))

    def tensors(self):
        return [t for _, _, v in self.entries for t in v]

    def take(self, kind, key):
        if self.pos == len(self.entries):  # another forward pass over the same batch
            self.pos = 0
        k, want, value = self.entries[self.pos]
        if k != kind or want != tuple(key):
            raise RuntimeError("geometry tape out of order: the model asked for %s%r where the tape holds %s%r"
                               % (kind, tuple(key), k, want))
        self.pos += 1
        return value


_tape = None


class replay_geometry:
    """Context manager: the layers inside take their geometry from ``tape`` (None = compute it)."""

    def __init__(self, tape):
        self.tape = tape

    def __enter__(self):
        global _tape
        self.prev, _tape = _tape, self.tape
        if self.tape is not None:
            self.tape.pos = 0
        return self.tape

    def __exit__(self, *exc):
        global _tape
        _tape = self.prev
        return False


def sampling_geometry(npoint, radius, nsample, xyz):
    """new_xyz (B,npoint,3) and the ball-query neighbour indices (B,npoint,nsample) of an SA layer
    (pointnet_util.py:39-46)."""
    if _tape is not None:
        return _tape.take("sample", (npoint, float(radius), nsample, tuple(xyz.shape)))
    new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
    return new_xyz, idx


def interpolation_geometry(xyz1, xyz2):
    """three nearest known points and their normalised inverse-distance weights (pointnet_util.py:299-303)."""
    if _tape is not None:
        return _tape.take("interp", (tuple(xyz1.shape), tuple(xyz2.shape)))
    dist, idx = three_nn(xyz1, xyz2)
    return idx, fp_weights(dist)


class _GroupConcat(torch.autograd.Function):
    """(B,m,ns,3+C) = [xyz[idx]-new_xyz | points[idx]] (xyz_first) or [points | xyz] (MSG)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, points, idx, xyz_first, use_xyz):
        b, n, _ = xyz.shape
        _, m, ns = idx.shape
        c = 0 if points is None else points.shape[2]
        w = (3 if use_xyz else 0) + c
        xyz_c, new_c = xyz.contiguous(), new_xyz.contiguous()
        pts_c = None if points is None else points.contiguous()
        # rows padded to a multiple of 4 floats: 16-byte aligned rows are what the GEMM's TMA
        # tensor maps need (3+C is odd for every SA layer); the result is the (.., :w) view
        ld = _pad4(w)
        buf = torch.empty((b, m, ns, ld), dtype=F32, device=xyz.device)
        call("pn2_group_concat_ld", b, n, m, ns, c, ptr(xyz_c, F32), ptr(new_c, F32),
             ptr(pts_c, F32, True), ptr(idx, I32), 1 if xyz_first else 0, 1 if use_xyz else 0,
             ptr(buf, F32), ld)
        ctx.save_for_backward(idx)
        ctx.cfg = (b, n, m, ns, c, xyz_first, use_xyz)
        return buf[..., :w] if ld != w else buf

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        b, n, m, ns, c, xyz_first, use_xyz = ctx.cfg
        g = g.contiguous()
        dev = g.device
        need_xyz, need_new, need_pts = ctx.needs_input_grad[0], ctx.needs_input_grad[1], \
            ctx.needs_input_grad[2] and c > 0
        g_pts = torch.empty((b, n, c), dtype=F32, device=dev) if need_pts else None
        g_xyz = torch.empty((b, n, 3), dtype=F32, device=dev) if (need_xyz and use_xyz) else None
        g_new = torch.empty((b, m, 3), dtype=F32, device=dev) if (need_new and use_xyz) else None
        if g_pts is not None or g_xyz is not None or g_new is not None:
            call("pn2_group_concat_grad", b, n, m, ns, c, ptr(g, F32), ptr(idx, I32),
                 1 if xyz_first else 0, 1 if use_xyz else 0, ptr(g_pts, F32, True),
                 ptr(g_xyz, F32, True), ptr(g_new, F32, True))
        return g_xyz, g_new, g_pts, None, None, None


def sample_and_group(npoint, radius, nsample, xyz, points, knn=False, use_xyz=True):
    """pointnet_util.py:18-60.  Returns new_xyz (B,npoint,3), new_points (B,npoint,nsample,3+C)
    with channel order [xyz, features], idx (B,npoint,nsample), grouped_xyz (centred)."""
    if knn:
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
        _, idx = knn_point(nsample, xyz, new_xyz)
    else:
        new_xyz, idx = sampling_geometry(npoint, radius, nsample, xyz)
    if points is not None and use_xyz:
        # one fused pass writes [xyz - centre | features]; grouped_xyz is its first 3 channels
        new_points = _GroupConcat.apply(xyz, new_xyz, points, idx, True, True)
        grouped_xyz = new_points[..., 0:3]
    else:
        grouped_xyz = _GroupConcat.apply(xyz, new_xyz, None, idx, True, True)
        new_points = group_point(points, idx) if points is not None else grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def sample_and_group_all(xyz, points, use_xyz=True):
    """pointnet_util.py:63-95: one group holding every point, centroid (0,0,0)."""
    b, n, _ = xyz.shape
    new_xyz = torch.zeros((b, 1, 3), dtype=F32, device=xyz.device)
    idx = torch.arange(n, dtype=I32, device=xyz.device).view(1, 1, n).repeat(b, 1, 1)
    grouped_xyz = xyz.reshape(b, 1, n, 3)
    if points is not None:
        new_points = torch.cat([xyz, points], dim=2) if use_xyz else points
        new_points = new_points.unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


POOL_MODES = {"avg": 1, "weighted_avg": 2, "max_and_avg": 3}


class _GroupPool(torch.autograd.Function):
    """feat (B,m,ns,N) -> (B,m,1,N) [avg, weighted_avg] or (B,m,1,2N) [max_and_avg = avg | max]."""

    @staticmethod
    def forward(ctx, feat, weights, mode, scope):
        b, m, ns, n = feat.shape
        x = feat.contiguous()
        g = b * m
        out = torch.empty((b, m, 1, 2 * n if mode == 3 else n), dtype=F32, device=x.device)
        arg = torch.empty((g, n), dtype=I32, device=x.device) if mode == 3 else None
        call("pn2_group_pool", g, ns, n, ptr(x, F32), ptr(weights, F32, True), mode, ptr(out, F32),
             ptr(arg, I32, True))
        if mode == 3 and tf_util.debug_capture is not None:
            tf_util.debug_capture[scope + "/argmax"] = arg
        ctx.cfg, ctx.weights, ctx.arg = (g, ns, n, mode), weights, arg
        return out

    @staticmethod
    def backward(ctx, d_out):
        g, ns, n, mode = ctx.cfg
        d = d_out.contiguous()
        dx = torch.empty((g * ns, n), dtype=F32, device=d.device)
        call("pn2_group_pool_grad", g, ns, n, ptr(d, F32), ptr(ctx.weights, F32, True),
             ptr(ctx.arg, I32, True), mode, ptr(dx, F32))
        return dx.view(d.shape[0], d.shape[1], ns, n), None, None, None


def _pool_weights(grouped_xyz):
    """exp(-5 |g|) normalised over nsample (pointnet_util.py:176-183); grouped_xyz (B,m,ns,>=3 strided)."""
    b, m, ns, _ = grouped_xyz.shape
    gx = grouped_xyz.detach()
    if gx.stride(3) != 1 or gx.stride(2) < 3 or gx.stride(1) != ns * gx.stride(2) or \
            gx.stride(0) != m * gx.stride(1):
        gx = gx.contiguous()
    w = torch.empty((b * m * ns,), dtype=F32, device=gx.device)
    import ctypes
    call("pn2_pool_weights", b * m, ns, ctypes.c_void_p(gx.data_ptr()), int(gx.stride(2)), ptr(w, F32))
    return w


def _conv_layers(prefix, k, widths, bn):
    layers = []
    for i, n in enumerate(widths):
        layers.append(tf_util.make_layer(prefix % i, k, n, bn, tf_util.relu))
        k = n
    return layers


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, group_all, is_training,
                       bn_decay, scope, bn=True, pooling="max", knn=False, use_xyz=True,
                       use_nchw=False):
    """PointNet Set Abstraction module (pointnet_util.py:98-216).

    Returns new_xyz (B,npoint,3), new_points (B,npoint,mlp[-1] or mlp2[-1]), idx."""
    training = tf_util._as_bool(is_training)
    with tf_util.variable_scope(scope):
        if group_all:
            nsample = xyz.shape[1]
            new_xyz, new_points, idx, grouped_xyz = sample_and_group_all(xyz, points, use_xyz)
        else:
            new_xyz, new_points, idx, grouped_xyz = sample_and_group(
                npoint, radius, nsample, xyz, points, knn, use_xyz)
        b, m, ns, k = new_points.shape
        layers = _conv_layers("conv%d", k, mlp, bn)
        x2d = new_points.reshape(b * m * ns, k)
        if pooling == "max":
            pooled = tf_util.mlp_chain(x2d, layers, training, bn_decay, pool_ns=ns)
            new_points = pooled.view(b, m, 1, mlp[-1])
        else:
            if pooling not in POOL_MODES:
                raise ValueError("unknown pooling %r" % (pooling,))
            feat = tf_util.mlp_chain(x2d, layers, training, bn_decay).view(b, m, ns, mlp[-1])
            w = _pool_weights(grouped_xyz) if pooling == "weighted_avg" else None
            new_points = _GroupPool.apply(feat, w, POOL_MODES[pooling], layers[-1].w.name[:-len("/weights")])
        if mlp2 is not None:
            kk = new_points.shape[-1]
            layers2 = _conv_layers("conv_post_%d", kk, mlp2, bn)
            new_points = tf_util.mlp_chain(new_points.reshape(b * m, kk), layers2, training,
                                           bn_decay).view(b, m, 1, mlp2[-1])
        return new_xyz, new_points.squeeze(2), idx


def pointnet_sa_module_msg(xyz, points, npoint, radius_list, nsample_list, mlp_list, is_training,
                           bn_decay, scope, bn=True, use_xyz=True, use_nchw=False):
    """SA module with multi-scale grouping (pointnet_util.py:219-282); per scale the channel
    order is [features, xyz] (:260), scales are concatenated in order."""
    training = tf_util._as_bool(is_training)
    with tf_util.variable_scope(scope):
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
        b, m, _ = new_xyz.shape
        outs = []
        for i, radius in enumerate(radius_list):
            ns = nsample_list[i]
            idx, _ = query_ball_point(radius, ns, xyz, new_xyz)
            if points is not None:
                grouped = _GroupConcat.apply(xyz, new_xyz, points, idx, False, bool(use_xyz))
            else:
                grouped = _GroupConcat.apply(xyz, new_xyz, None, idx, True, True)
            k = grouped.shape[-1]
            layers = []
            for j, n in enumerate(mlp_list[i]):
                layers.append(tf_util.make_layer("conv%d_%d" % (i, j), k, n, bn, tf_util.relu))
                k = n
            pooled = tf_util.mlp_chain(grouped.reshape(b * m * ns, grouped.shape[-1]), layers,
                                       training, bn_decay, pool_ns=ns)
            outs.append(pooled.view(b, m, k))
        return new_xyz, torch.cat(outs, dim=-1)


class _InterpConcat(torch.autograd.Function):
    """X0 (B*n, C2+C1) = [three_interpolate(points2, idx, w) | points1], written in place."""

    @staticmethod
    def forward(ctx, points2, points1, idx, weight):
        b, m, c2 = points2.shape
        n = idx.shape[1]
        c1 = 0 if points1 is None else points1.shape[2]
        w = c2 + c1
        p2 = points2.contiguous()
        ld = _pad4(w)  # 16-byte aligned rows for the GEMM's TMA tensor maps
        buf = torch.empty((b * n, ld), dtype=F32, device=points2.device)
        call("pn2_three_interpolate_ld", b, m, c2, n, ptr(p2, F32), ptr(idx, I32), ptr(weight, F32),
             ptr(buf, F32), ld)
        if c1:
            p1 = points1.contiguous()
            call("pn2_copy_cols", b * n, c1, ptr(p1, F32), c1,
                 _ffi_offset(buf, c2), ld, 0)
        ctx.save_for_backward(idx, weight)
        ctx.cfg = (b, m, c2, n, c1)
        return buf[:, :w] if ld != w else buf

    @staticmethod
    def backward(ctx, g):
        idx, weight = ctx.saved_tensors
        b, m, c2, n, c1 = ctx.cfg
        g = g.contiguous()
        w = c2 + c1
        g2 = g1 = None
        if ctx.needs_input_grad[0]:
            g2 = torch.empty((b, m, c2), dtype=F32, device=g.device)
            call("pn2_three_interpolate_grad_ld", b, n, c2, m, ptr(g, F32), w, ptr(idx, I32),
                 ptr(weight, F32), ptr(g2, F32))
        if c1 and ctx.needs_input_grad[1]:
            g1 = torch.empty((b, n, c1), dtype=F32, device=g.device)
            call("pn2_copy_cols", b * n, c1, _ffi_offset(g, c2), w, ptr(g1, F32), c1, 0)
        return g2, g1, None, None


def _ffi_offset(t, col):
    """device address of column ``col`` of the first row of a contiguous 2-D fp32 tensor"""
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + 4 * col)


def fp_weights(dist):
    """pointnet_util.py:300-303: inverse (squared) distance weights, floored at 1e-10."""
    dist = dist.contiguous()
    w = torch.empty_like(dist)
    call("pn2_fp_weights", dist.numel() // 3, ptr(dist, F32), ptr(w, F32))
    return w


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope, bn=True):
    """PointNet Feature Propagation module (pointnet_util.py:285-326).
    xyz1 (B,n1,3) dense, xyz2 (B,n2,3) sparse, points1 (B,n1,C1) or None, points2 (B,n2,C2)
    -> (B,n1,mlp[-1]);  concat order [interpolated, points1] (:307-309)."""
    training = tf_util._as_bool(is_training)
    with tf_util.variable_scope(scope):
        idx, weight = interpolation_geometry(xyz1, xyz2)
        x0 = _InterpConcat.apply(points2, points1, idx, weight)
        b, n1, _ = xyz1.shape
        layers = _conv_layers("conv_%d", x0.shape[1], mlp, bn)
        # concat order [interpolated | points1]: when points1 needs no gradient (the network input at FP4) only the
        # interpolated columns of the input gradient are read
        dx_cols = None
        if points1 is not None and not points1.requires_grad:
            dx_cols = (0, points2.shape[2])
        out = tf_util.mlp_chain(x0, layers, training, bn_decay, dx_cols=dx_cols)
        return out.view(b, n1, mlp[-1])
