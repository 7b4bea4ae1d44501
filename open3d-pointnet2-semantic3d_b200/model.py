"""Semantic-segmentation PointNet++ (SSG): the reference's model.py on the sm_100a engine.

  get_placeholders(num_point, hyperparams)                         model.py:12-19
  get_model(point_cloud, is_training, num_class, hyperparams,
            bn_decay=None) -> (logits (B,N,num_class), end_points) model.py:22-148
  get_loss(pred, label, smpw, end_points) -> scalar loss           model.py:152-161

The call sites into pointnet_sa_module / pointnet_fp_module / tf_util.conv1d / dropout are the
reference's, keyword for keyword (scopes layer1..4, fa_layer1..4, fc1, dp1, fc2), so the
variable names match the reference's checkpoints.
"""
import collections

import torch

from ._ffi import F32, F64, I32, call, ptr
from .util import tf_util
from .util.pointnet_util import pointnet_fp_module, pointnet_sa_module

Placeholder = collections.namedtuple("Placeholder", "dtype shape")

SA_MLPS = ([32, 32, 64], [64, 64, 128], [128, 128, 256], [256, 256, 512])
FP_MLPS = ([256, 256], [256, 256], [256, 128], [128, 128, 128])


def get_placeholders(num_point, hyperparams):
    """Shapes/dtypes of the three inputs (there is no graph to feed; model.py:12-19)."""
    feature_size = 3 * int(hyperparams["use_color"])
    return (Placeholder(torch.float32, (None, num_point, 3 + feature_size)),
            Placeholder(torch.int32, (None, num_point)),
            Placeholder(torch.float32, (None, num_point)))


def get_model(point_cloud, is_training, num_class, hyperparams, bn_decay=None):
    """point_cloud (B,N,3+feat) CUDA float32 -> logits (B,N,num_class), end_points dict."""
    end_points = {}
    if hyperparams["use_color"]:
        feature_size = 3 * int(hyperparams["use_color"])
        l0_xyz = point_cloud[:, :, 0:3].contiguous()
        l0_points = point_cloud[:, :, 3:3 + feature_size].contiguous()
    else:
        l0_xyz, l0_points = point_cloud, None
    end_points["l0_xyz"] = l0_xyz

    xyz, pts = [l0_xyz], [l0_points]
    for l in (1, 2, 3, 4):
        nx, npts, _ = pointnet_sa_module(
            xyz[-1], pts[-1],
            npoint=hyperparams["l%d_npoint" % l],
            radius=hyperparams["l%d_radius" % l],
            nsample=hyperparams["l%d_nsample" % l],
            mlp=SA_MLPS[l - 1], mlp2=None, group_all=False,
            is_training=is_training, bn_decay=bn_decay, scope="layer%d" % l)
        xyz.append(nx)
        pts.append(npts)

    # feature propagation: (l3<-l4), (l2<-l3), (l1<-l2), (l0<-l1)
    up = pts[4]
    for l, lo in zip((1, 2, 3, 4), (3, 2, 1, 0)):
        up = pointnet_fp_module(xyz[lo], xyz[lo + 1], pts[lo], up, FP_MLPS[l - 1], is_training,
                                bn_decay, scope="fa_layer%d" % l)

    net = tf_util.conv1d(up, 128, 1, padding="VALID", bn=True, is_training=is_training,
                         scope="fc1", bn_decay=bn_decay)
    end_points["feats"] = net
    net = tf_util.dropout(net, keep_prob=0.5, is_training=is_training, scope="dp1")
    net = tf_util.conv1d(net, num_class, 1, padding="VALID", activation_fn=None, scope="fc2")
    return net, end_points


def get_geometry(point_cloud, hyperparams):
    """The weight-independent part of get_model for one batch -- per SA layer the sampled centres and their
    ball-query neighbours, per FP layer the three nearest known points and their weights -- as a GeometryTape
    in get_model's call order (util/pointnet_util.py).  `with replay_geometry(tape): get_model(...)` then runs
    only the dense stage.  Same kernels, same results as computing them inside the layers."""
    from .util import pointnet_util as pu
    tape = pu.GeometryTape()
    with torch.no_grad(), pu.replay_geometry(None):
        xyz = [point_cloud[:, :, 0:3].contiguous() if hyperparams["use_color"] else point_cloud]
        for l in (1, 2, 3, 4):
            npoint, radius, nsample = (hyperparams["l%d_%s" % (l, k)] for k in ("npoint", "radius", "nsample"))
            new_xyz, idx = pu.sampling_geometry(npoint, radius, nsample, xyz[-1])
            tape.add("sample", (npoint, float(radius), nsample, tuple(xyz[-1].shape)), (new_xyz, idx))
            xyz.append(new_xyz)
        for lo in (3, 2, 1, 0):
            idx, weight = pu.interpolation_geometry(xyz[lo], xyz[lo + 1])
            tape.add("interp", (tuple(xyz[lo].shape), tuple(xyz[lo + 1].shape)), (idx, weight))
    return tape


class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, smpw):
        rows, c = pred.numel() // pred.shape[-1], pred.shape[-1]
        pred = pred.contiguous()
        acc = tf_util.zero_arena.take(2, pred.device)
        loss = torch.empty((), dtype=F32, device=pred.device)
        call("pn2_softmax_ce_reduce", rows, c, ptr(pred, F32), ptr(label, I32),
             ptr(smpw, F32, True), ptr(acc, F64))
        call("pn2_softmax_ce_grad", rows, c, ptr(pred, F32), ptr(label, I32), ptr(smpw, F32, True),
             ptr(acc, F64), 1.0, None, ptr(loss, F32), None)
        ctx.save_for_backward(pred, label, smpw, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, label, smpw, acc = ctx.saved_tensors
        rows, c = pred.numel() // pred.shape[-1], pred.shape[-1]
        d = torch.empty_like(pred)
        # the upstream gradient g (0-dim) is read on the device: no host synchronisation
        gc = g.contiguous()  # named: the buffer must outlive the launch
        call("pn2_softmax_ce_grad", rows, c, ptr(pred, F32), ptr(label, I32), ptr(smpw, F32, True),
             ptr(acc, F64), 1.0, ptr(gc, F32), None, ptr(d, F32))
        return d, None, None


def get_loss(pred, label, smpw, end_points=None):
    """Weighted sparse softmax cross entropy, tf.losses SUM_BY_NONZERO_WEIGHTS (model.py:152-161).
    pred (B,N,C) float32, label (B,N) int32, smpw (B,N) float32 -> 0-dim loss."""
    label = label.to(torch.int32).contiguous()
    smpw = None if smpw is None else smpw.to(torch.float32).contiguous()
    return _SoftmaxCE.apply(pred, label, smpw)
