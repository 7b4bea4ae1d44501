"""ctypes/numpy front end of the CPU oracle (oracle/pn2_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  The product package never
imports this module (tests/test_no_oracle_in_product.py enforces it).

Every function mirrors one reference op (reference file:line in pn2_oracle.c) and
takes/returns C-contiguous numpy arrays (float32 / int32), exactly the dense
row-major layout the reference kernels use.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_ci = ctypes.c_int


def build(force=False):
    """Compile liborc.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "pn2_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    if a.dtype == np.float32:
        return a.ctypes.data_as(_f)
    return a.ctypes.data_as(_i)


def farthest_point_sample(npoint, inp, threads=0):
    inp = _f32(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    lib().orc_fps(_ci(b), _ci(n), _ci(npoint), _p(inp), _p(out), _ci(threads))
    return out


def gather_point(inp, idx):
    inp, idx = _f32(inp), _i32(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.empty((b, m, 3), np.float32)
    lib().orc_gather_point(_ci(b), _ci(n), _ci(m), _p(inp), _p(idx), _p(out))
    return out


def gather_point_grad(inp_shape, idx, out_g):
    idx, out_g = _i32(idx), _f32(out_g)
    b, n, _ = inp_shape
    m = idx.shape[1]
    g = np.empty((b, n, 3), np.float32)
    lib().orc_gather_point_grad(_ci(b), _ci(n), _ci(m), _p(out_g), _p(idx), _p(g))
    return g


def query_ball_point(radius, nsample, xyz1, xyz2, threads=0):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.empty((b, m, nsample), np.int32)
    cnt = np.empty((b, m), np.int32)
    lib().orc_query_ball_point(_ci(b), _ci(n), _ci(m), ctypes.c_float(radius), _ci(nsample),
                               _p(xyz1), _p(xyz2), _p(idx), _p(cnt), _ci(threads))
    return idx, cnt


def group_point(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    lib().orc_group_point(_ci(b), _ci(n), _ci(c), _ci(m), _ci(ns), _p(points), _p(idx), _p(out))
    return out


def group_point_grad(points_shape, idx, grad_out):
    idx, grad_out = _i32(idx), _f32(grad_out)
    b, n, c = points_shape
    _, m, ns = idx.shape
    g = np.empty((b, n, c), np.float32)
    lib().orc_group_point_grad(_ci(b), _ci(n), _ci(c), _ci(m), _ci(ns), _p(grad_out), _p(idx),
                               _p(g))
    return g


def three_nn(xyz1, xyz2, threads=0):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().orc_three_nn(_ci(b), _ci(n), _ci(m), _p(xyz1), _p(xyz2), _p(dist), _p(idx),
                       _ci(threads))
    return dist, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    lib().orc_three_interpolate(_ci(b), _ci(m), _ci(c), _ci(n), _p(points), _p(idx), _p(weight),
                                _p(out))
    return out


def three_interpolate_grad(points_shape, idx, weight, grad_out):
    idx, weight, grad_out = _i32(idx), _f32(weight), _f32(grad_out)
    b, m, c = points_shape
    n = idx.shape[1]
    g = np.empty((b, m, c), np.float32)
    lib().orc_three_interpolate_grad(_ci(b), _ci(n), _ci(c), _ci(m), _p(grad_out), _p(idx),
                                     _p(weight), _p(g))
    return g


def interpolate_label_with_color(sparse_points, sparse_labels, dense_points, knn, threads=0):
    """tf_interpolate.cpp:71-115 -> (dense_labels (N,) int32, dense_colors (N,3) uint8)."""
    sp, sl, dp = _f32(sparse_points), _i32(sparse_labels), _f32(dense_points)
    ns, nd = sp.shape[0], dp.shape[0]
    labels = np.empty((nd,), np.int32)
    colors = np.empty((nd, 3), np.uint8)
    lib().orc_interpolate_label_with_color(
        _ci(ns), _ci(nd), _p(sp), _p(sl), _p(dp), _p(labels),
        colors.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), _ci(int(knn)), _ci(threads))
    return labels, colors


def select_top_k(k, dist):
    dist = _f32(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    lib().orc_selection_sort(_ci(b), _ci(n), _ci(m), _ci(k), _p(dist), _p(outi), _p(out))
    return outi, out


def knn_point(k, xyz1, xyz2):
    """tf_grouping.py:64-89: (val (b,m,k) squared distances, idx (b,m,k)) -- the first k entries of the
    reference's selection sort over the fp32 distance rows."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    val = np.empty((b, m, k), np.float32)
    idx = np.empty((b, m, k), np.int32)
    lib().orc_knn_point(_ci(b), _ci(n), _ci(c), _ci(m), _ci(k), _p(xyz1), _p(xyz2), _p(val), _p(idx))
    return val, idx


def cumsum(inp):
    """Row-wise inclusive prefix sum in the reference's rounding order (tf_sampling.cu:7-92)."""
    inp = _f32(inp)
    b, n = inp.shape
    out = np.empty((b, n), np.float32)
    lib().orc_cumsum(_ci(b), _ci(n), _p(inp), _p(out))
    return out


def prob_sample(inp, inpr):
    inp, inpr = _f32(inp), _f32(inpr)
    b, n = inp.shape
    m = inpr.shape[1]
    out = np.empty((b, m), np.int32)
    lib().orc_prob_sample(_ci(b), _ci(n), _ci(m), _p(inp), _p(inpr), _p(out))
    return out
