"""Grouping ops: same names / argument order as the reference's tf_ops/tf_grouping.py.

  query_ball_point(radius, nsample, xyz1, xyz2)  tf_grouping.py:13-25 (no gradient, :28)
  group_point(points, idx)                       tf_grouping.py:46-54 (gradient :57-61)
  select_top_k(k, dist)                          tf_grouping.py:31-40
  knn_point(k, xyz1, xyz2)                       tf_grouping.py:64-89

Validation texts follow tf_grouping.cpp:80-105, 187-200, 142-151.
"""
import os

import torch

from .._ffi import F32, I32, call, ptr


def _need(cond, msg):
    if not cond:
        raise ValueError(msg)


def query_ball_point(radius, nsample, xyz1, xyz2):
    """xyz1 (B,n,3) data, xyz2 (B,m,3) queries -> idx (B,m,nsample) int32, pts_cnt (B,m) int32.

    First ``nsample`` points (ascending index) with max(sqrt(d2),1e-20) < radius; shorter rows
    are padded with the first hit; rows without any hit are zeros (reference: uninitialised)."""
    _need(float(radius) > 0, "QueryBallPoint expects positive radius")
    _need(int(nsample) > 0, "QueryBallPoint expects positive nsample")
    _need(xyz1.dim() == 3 and xyz1.shape[2] == 3,
          "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.")
    _need(xyz2.dim() == 3 and xyz2.shape[2] == 3,
          "QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    xyz1 = xyz1.detach().contiguous()
    xyz2 = xyz2.detach().contiguous()
    idx = torch.empty((b, m, int(nsample)), dtype=I32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=I32, device=xyz1.device)
    if _GRID and n >= 4096:  # hashed uniform grid: 2x (n=8192) to 14x (n=262144) faster, bit-identical
        ws, nbytes = _grid_workspace(b, n, xyz1.device)
        call("pn2_query_ball_point_grid", b, n, m, float(radius), int(nsample), ptr(xyz1, F32),
             ptr(xyz2, F32), ptr(idx, I32), ptr(cnt, I32), ptr(ws, F32), nbytes)
        return idx, cnt
    call("pn2_query_ball_point", b, n, m, float(radius), int(nsample), ptr(xyz1, F32),
         ptr(xyz2, F32), ptr(idx, I32), ptr(cnt, I32))
    return idx, cnt


# Large clouds go through the hashed-grid kernel (profiles/ab_ops_r02.json: 0.096 vs 0.184 ms at the SA1
# size of semantic.json, 3-14x at 65 k-262 k points; slower below ~4 k points, where the TMA brute-force
# kernels stay).  PN2_BALL_GRID=0 keeps the brute-force kernels everywhere.
_GRID = os.environ.get("PN2_BALL_GRID", "1") != "0"
_grid_ws = {}


def _grid_workspace(b, n, device):
    """Caller-owned scratch of pn2_query_ball_point_grid, one per (b, n, device)."""
    key = (b, n, str(device))
    hit = _grid_ws.get(key)
    if hit is None:
        from .._ffi import lib
        nbytes = int(lib().pn2_ball_grid_workspace_bytes(b, n))
        hit = (torch.empty((nbytes + 15) // 16 * 4, dtype=F32, device=device), nbytes)
        _grid_ws[key] = hit
    return hit


class _GroupPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        b, n, c = points.shape
        _, m, ns = idx.shape
        points_c = points.contiguous()
        out = torch.empty((b, m, ns, c), dtype=F32, device=points.device)
        call("pn2_group_point", b, n, c, m, ns, ptr(points_c, F32), ptr(idx, I32), ptr(out, F32))
        ctx.save_for_backward(idx)
        ctx.dims = (b, n, c, m, ns)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        b, n, c, m, ns = ctx.dims
        grad_out = grad_out.contiguous()
        g = torch.empty((b, n, c), dtype=F32, device=grad_out.device)
        call("pn2_group_point_grad", b, n, c, m, ns, ptr(grad_out, F32), ptr(idx, I32), ptr(g, F32))
        return g, None


def group_point(points, idx):
    """points (B,n,C), idx (B,m,ns) int32 -> (B,m,ns,C); differentiable w.r.t. points."""
    _need(points.dim() == 3, "GroupPoint expects (batch_size, num_points, channel) points shape")
    _need(idx.dim() == 3 and idx.shape[0] == points.shape[0],
          "GroupPoint expects (batch_size, npoints, nsample) idx shape")
    return _GroupPoint.apply(points, idx.contiguous())


def group_point_grad(points, idx, grad_out):
    """Explicit gradient op (reference: grouping_module.group_point_grad)."""
    b, n, c = points.shape
    _, m, ns = idx.shape
    _need(grad_out.dim() == 4 and tuple(grad_out.shape) == (b, m, ns, c),
          "GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape")
    g = torch.empty((b, n, c), dtype=F32, device=points.device)
    go, ii = grad_out.contiguous(), idx.contiguous()
    call("pn2_group_point_grad", b, n, c, m, ns, ptr(go, F32), ptr(ii, I32), ptr(g, F32))
    return g


def select_top_k(k, dist):
    """dist (B,m,n) -> (idx (B,m,n) int32, dist_out (B,m,n)): the reference's partial selection sort
    (tf_grouping.cu:95-136) bit for bit -- the k smallest ascending in [..., :k] (ties in the reference's
    order), the remainder of each row in the order its swaps leave it."""
    _need(int(k) > 0, "SelectionSort expects positive k")
    _need(dist.dim() == 3, "SelectionSort expects (b,m,n) dist shape.")
    b, m, n = dist.shape
    dist = dist.detach().contiguous()
    outi = torch.empty((b, m, n), dtype=I32, device=dist.device)
    out = torch.empty((b, m, n), dtype=F32, device=dist.device)
    call("pn2_selection_sort", b, n, m, int(k), ptr(dist, F32), ptr(outi, I32), ptr(out, F32))
    return outi, out


def knn_point(k, xyz1, xyz2):
    """k nearest data points of every query (tf_grouping.py:64-89).  xyz1 (B,n,c) data, xyz2 (B,m,c)
    queries -> (val (B,m,k) squared distances, idx (B,m,k) int32).  One fused kernel: distances on the
    fly + the reference's selection, no (b,m,n) tensor (csrc/pn2_grouping.cu knn_point_kernel)."""
    k = int(k)
    _need(k > 0, "SelectionSort expects positive k")
    _need(xyz1.dim() == 3 and xyz2.dim() == 3 and xyz1.shape[0] == xyz2.shape[0]
          and xyz1.shape[2] == xyz2.shape[2], "knn_point expects (b,n,c) xyz1 and (b,m,c) xyz2")
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    _need(k <= n, "knn_point: k must not exceed the number of data points")
    x1, x2 = xyz1.detach().contiguous(), xyz2.detach().contiguous()
    val = torch.empty((b, m, k), dtype=F32, device=x1.device)
    idx = torch.empty((b, m, k), dtype=I32, device=x1.device)
    call("pn2_knn_point", b, n, c, m, k, ptr(x1, F32), ptr(x2, F32), ptr(val, F32), ptr(idx, I32))
    return val, idx
