"""Interpolation ops: same names / argument order as the reference's tf_ops/tf_interpolate.py.

  interpolate_label_with_color(sparse_points, sparse_labels, dense_points, knn)
                                          tf_interpolate.py:26-43 (inference post-processing)
  three_nn(xyz1, xyz2)                    tf_interpolate.py:13-22 (no gradient, :25)
  three_interpolate(points, idx, weight)  tf_interpolate.py:50-59 (gradient :62-71)

The reference runs both on the CPU (KD-tree / scalar loops, tf_interpolate.cpp:213-243,
307-330, 397-421); here they are GPU kernels with the same results
(csrc/pn2_interpolate.cu).  Validation texts follow tf_interpolate.cpp:254-266, 343-364.
"""
import torch

from .._ffi import F32, I32, call, ptr


def _need(cond, msg):
    if not cond:
        raise ValueError(msg)


def interpolate_label_with_color(sparse_points, sparse_labels, dense_points, knn):
    """sparse_points (Ns,3) float32, sparse_labels (Ns,) int32, dense_points (Nd,3) float32,
    knn int -> (dense_labels (Nd,) int32, dense_colors (Nd,3) uint8).

    Label vote among the knn nearest sparse points of every dense point (nearest first; the
    label whose running count first becomes the largest wins) plus the reference's colour table
    (tf_interpolate.cpp:46-48, 71-115).  Validation texts: tf_interpolate.cpp:125-160."""
    _need(sparse_points.dim() == 2 and sparse_points.shape[1] == 3,
          "sparse_points must be: (num_sparse_points, 3)")
    ns = sparse_points.shape[0]
    _need(sparse_labels.dim() == 1 and sparse_labels.shape[0] == ns,
          "sparse_labels must be: (num_sparse_points, 3)")  # sic: the reference's message
    _need(dense_points.dim() == 2 and dense_points.shape[1] == 3,
          "dense_points must be: (num_dense_points, 3)")
    _need(isinstance(knn, int) or (isinstance(knn, torch.Tensor) and knn.dim() == 0),
          "knn must be an int scalar")
    knn = int(knn)
    _need(knn > 0, "knn must be an int scalar")
    nd = dense_points.shape[0]
    sp = sparse_points.detach().contiguous()
    sl = sparse_labels.contiguous()
    dp = dense_points.detach().contiguous()
    labels = torch.empty((nd,), dtype=I32, device=dp.device)
    colors = torch.empty((nd, 3), dtype=torch.uint8, device=dp.device)
    call("pn2_interpolate_label_with_color", ns, nd, ptr(sp, F32), ptr(sl, I32), ptr(dp, F32),
         ptr(labels, I32), ptr(colors, torch.uint8), knn)
    return labels, colors


def three_nn(xyz1, xyz2):
    """xyz1 (b,n,3) queries, xyz2 (b,m,3) known -> dist (b,n,3) SQUARED, idx (b,n,3) int32."""
    _need(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ThreeNN expects (b,n,3) xyz1 shape.")
    _need(xyz2.dim() == 3 and xyz2.shape[2] == 3, "ThreeNN expects (b,m,3) xyz2 shape.")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    _need(m >= 3, "ThreeNN needs at least 3 known points")
    xyz1 = xyz1.detach().contiguous()
    xyz2 = xyz2.detach().contiguous()
    dist = torch.empty((b, n, 3), dtype=F32, device=xyz1.device)
    idx = torch.empty((b, n, 3), dtype=I32, device=xyz1.device)
    call("pn2_three_nn", b, n, m, ptr(xyz1, F32), ptr(xyz2, F32), ptr(dist, F32), ptr(idx, I32))
    return dist, idx


class _ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, weight):
        b, m, c = points.shape
        n = idx.shape[1]
        points_c = points.contiguous()
        out = torch.empty((b, n, c), dtype=F32, device=points.device)
        call("pn2_three_interpolate", b, m, c, n, ptr(points_c, F32), ptr(idx, I32),
             ptr(weight, F32), ptr(out, F32))
        ctx.save_for_backward(idx, weight)
        ctx.dims = (b, m, c, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        b, m, c, n = ctx.dims
        grad_out = grad_out.contiguous()
        g = torch.empty((b, m, c), dtype=F32, device=grad_out.device)
        call("pn2_three_interpolate_grad", b, n, c, m, ptr(grad_out, F32), ptr(idx, I32),
             ptr(weight, F32), ptr(g, F32))
        return g, None, None


def three_interpolate(points, idx, weight):
    """points (b,m,c), idx (b,n,3) int32, weight (b,n,3) -> (b,n,c); grad w.r.t. points only."""
    _need(points.dim() == 3, "ThreeInterpolate expects (b,m,c) points shape")
    b = points.shape[0]
    _need(idx.dim() == 3 and idx.shape[0] == b and idx.shape[2] == 3,
          "ThreeInterpolate expects (b,n,3) idx shape")
    _need(weight.dim() == 3 and tuple(weight.shape) == (b, idx.shape[1], 3),
          "ThreeInterpolate expects (b,n,3) weight shape")
    return _ThreeInterpolate.apply(points, idx.contiguous(), weight.detach().contiguous())


def three_interpolate_grad(points, idx, weight, grad_out):
    """Explicit gradient op (reference: interpolate_module.three_interpolate_grad)."""
    b, m, c = points.shape
    n = idx.shape[1]
    _need(grad_out.dim() == 3 and tuple(grad_out.shape) == (b, n, c),
          "ThreeInterpolateGrad expects (b,n,c) grad_out shape")
    g = torch.empty((b, m, c), dtype=F32, device=points.device)
    go, ii, ww = grad_out.contiguous(), idx.contiguous(), weight.contiguous()
    call("pn2_three_interpolate_grad", b, n, c, m, ptr(go, F32), ptr(ii, I32), ptr(ww, F32),
         ptr(g, F32))
    return g
