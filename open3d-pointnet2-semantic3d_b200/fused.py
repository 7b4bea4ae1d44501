"""Whole layers through ONE native call: ctypes binding of pn2_sa_forward/backward and
pn2_fp_forward/backward (include/pn2_b200.h "group 2b", csrc/pn2_layers.cu).

The host side of a layer -- the op sequence of pointnet_sa_module (util/pointnet_util.py:98-216) /
pointnet_fp_module (:285-326) and its reverse -- runs in C++ here instead of Python; the layer variables are
the same ``tf_util`` variables (created under the same scope names), all intermediates live in one
caller-owned workspace tensor.  ``util/pointnet_util.py`` remains the reference-shaped surface; this module
is the thin doorway for callers that want a layer per call (and the test vehicle of those entry points).
"""
import ctypes

import torch

from ._ffi import F32, I32, call, lib, ptr
from .util import tf_util

MAX_LAYERS = 8


class ConvLayer(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int), ("N", ctypes.c_int), ("bn", ctypes.c_int), ("relu", ctypes.c_int),
                ("rank4", ctypes.c_int)] + [(n, ctypes.c_void_p) for n in (
                    "W", "bias", "gamma", "beta", "moving_mean", "moving_var", "dW", "dbias", "dgamma", "dbeta")]


class SaConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("b", "n", "c", "npoint", "nsample", "nlayers", "is_training")] + \
               [(n, ctypes.c_float) for n in ("radius", "bn_eps", "bn_decay")]


class FpConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("b", "n1", "n2", "c1", "c2", "nlayers", "is_training")] + \
               [(n, ctypes.c_float) for n in ("bn_eps", "bn_decay")]


def _layer_table(scope_fmt, k, widths, bn, with_grads):
    """tf_util variables of a conv chain (created like pointnet_util does) -> (ctypes array, LayerSpecs)."""
    specs = []
    for i, n in enumerate(widths):
        specs.append(tf_util.make_layer(scope_fmt % i, k, n, bn, tf_util.relu))
        k = n
    arr = (ConvLayer * len(specs))()
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    for rec, L in zip(arr, specs):
        rec.K, rec.N, rec.bn, rec.relu, rec.rank4 = L.k, L.n, int(L.bn), int(L.relu), int(L.rank4)
        rec.W, rec.bias = L.w.data.data_ptr(), L.b.data.data_ptr()
        if L.bn:
            rec.gamma, rec.beta = L.gamma.data.data_ptr(), L.beta.data.data_ptr()
            rec.moving_mean, rec.moving_var = L.mm.data.data_ptr(), L.mv.data.data_ptr()
        if with_grads:
            rec.dW, rec.dbias = L.w.ensure_grad().data_ptr(), L.b.ensure_grad().data_ptr()
            if L.bn:
                rec.dgamma, rec.dbeta = p(L.gamma.ensure_grad()), p(L.beta.ensure_grad())
    return arr, specs


def _workspace(nbytes, device):
    if nbytes < 0:
        raise ValueError("layer configuration rejected by the native planner")
    return torch.empty((nbytes + 3) // 4, dtype=F32, device=device)


class SetAbstraction:
    """pointnet_sa_module (ball query, max pooling, no mlp2) as one native call per direction."""

    def __init__(self, scope, c, npoint, radius, nsample, mlp, bn=True):
        self.scope, self.c, self.npoint, self.radius, self.nsample, self.mlp, self.bn = \
            scope, int(c), int(npoint), float(radius), int(nsample), list(mlp), bool(bn)

    def forward(self, xyz, points, is_training=True, bn_decay=0.9):
        b, n, _ = xyz.shape
        with tf_util.variable_scope(self.scope):
            self.layers, _ = _layer_table("conv%d", 3 + self.c, self.mlp, self.bn, bool(is_training))
        self.cfg = SaConfig(b, n, self.c, self.npoint, self.nsample, len(self.mlp), int(bool(is_training)),
                            self.radius, tf_util.BN_EPS, float(bn_decay))
        nbytes = int(lib().pn2_sa_workspace_bytes(ctypes.byref(self.cfg), self.layers))
        self.ws = _workspace(nbytes, xyz.device)
        xyz = xyz.contiguous()
        pts = None if points is None else points.contiguous()
        new_xyz = torch.empty((b, self.npoint, 3), dtype=F32, device=xyz.device)
        new_points = torch.empty((b, self.npoint, self.mlp[-1]), dtype=F32, device=xyz.device)
        idx = torch.empty((b, self.npoint, self.nsample), dtype=I32, device=xyz.device)
        call("pn2_sa_forward", ctypes.byref(self.cfg), self.layers, ptr(xyz, F32), ptr(pts, F32, True),
             ptr(new_xyz, F32), ptr(new_points, F32), ptr(idx, I32), ptr(self.ws, F32), nbytes)
        self.idx, self.nbytes, self.shape = idx, nbytes, (b, n)
        return new_xyz, new_points, idx

    def backward(self, d_new_points):
        b, n = self.shape
        d_points = torch.empty((b, n, self.c), dtype=F32, device=d_new_points.device) if self.c else None
        g = d_new_points.contiguous()
        call("pn2_sa_backward", ctypes.byref(self.cfg), self.layers, ptr(g, F32), ptr(self.idx, I32),
             ptr(d_points, F32, True), ptr(self.ws, F32), self.nbytes)
        return d_points


class FeaturePropagation:
    """pointnet_fp_module as one native call per direction."""

    def __init__(self, scope, c1, c2, mlp, bn=True):
        self.scope, self.c1, self.c2, self.mlp, self.bn = scope, int(c1), int(c2), list(mlp), bool(bn)

    def forward(self, xyz1, xyz2, points1, points2, is_training=True, bn_decay=0.9):
        b, n1, _ = xyz1.shape
        n2 = xyz2.shape[1]
        with tf_util.variable_scope(self.scope):
            self.layers, _ = _layer_table("conv_%d", self.c2 + self.c1, self.mlp, self.bn, bool(is_training))
        self.cfg = FpConfig(b, n1, n2, self.c1, self.c2, len(self.mlp), int(bool(is_training)), tf_util.BN_EPS,
                            float(bn_decay))
        nbytes = int(lib().pn2_fp_workspace_bytes(ctypes.byref(self.cfg), self.layers))
        self.ws = _workspace(nbytes, xyz1.device)
        p1 = None if points1 is None else points1.contiguous()
        out = torch.empty((b, n1, self.mlp[-1]), dtype=F32, device=xyz1.device)
        call("pn2_fp_forward", ctypes.byref(self.cfg), self.layers, ptr(xyz1.contiguous(), F32),
             ptr(xyz2.contiguous(), F32), ptr(p1, F32, True), ptr(points2.contiguous(), F32), ptr(out, F32),
             ptr(self.ws, F32), nbytes)
        self.nbytes, self.shape = nbytes, (b, n1, n2)
        return out

    def backward(self, d_out):
        b, n1, n2 = self.shape
        dev = d_out.device
        d1 = torch.empty((b, n1, self.c1), dtype=F32, device=dev) if self.c1 else None
        d2 = torch.empty((b, n2, self.c2), dtype=F32, device=dev)
        g = d_out.contiguous()
        call("pn2_fp_backward", ctypes.byref(self.cfg), self.layers, ptr(g, F32), ptr(d1, F32, True), ptr(d2, F32),
             ptr(self.ws, F32), self.nbytes)
        return d1, d2
