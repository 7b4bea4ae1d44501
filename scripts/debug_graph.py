"""Diagnostic: which entry point invalidates CUDA-graph stream capture?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200 import _ffi as ffi
from pn2_b200.tf_ops import tf_sampling as ts, tf_grouping as tg, tf_interpolate as ti
p = ffi.ptr
dev = "cuda"
xyz = torch.rand(4, 2048, 3, device=dev)
A = torch.randn(4096, 128, device=dev); W = torch.randn(128, 128, device=dev) * .1
Y = torch.empty(4096, 128, device=dev); dX = torch.empty(4096, 128, device=dev); dW = torch.zeros(128, 128, device=dev)
A2 = torch.randn(131072, 64, device=dev); dY2 = torch.randn(131072, 64, device=dev); dW2 = torch.zeros(64, 64, device=dev)
A3 = torch.randn(8192, 32, device=dev); dY3 = torch.randn(8192, 32, device=dev); dW3 = torch.zeros(32, 32, device=dev)
stats = torch.zeros(256, dtype=torch.float64, device=dev)
ws = torch.empty(int(ffi.lib().pn2_linear_workspace_bytes(128, 128)) // 4, device=dev)
sc = torch.ones(128, device=dev); sh = torch.zeros(128, device=dev); saved = torch.zeros(256, device=dev)
red = torch.zeros(256, dtype=torch.float64, device=dev)

ops = {
    "torch.zeros f64": lambda: torch.zeros(64, dtype=torch.float64, device=dev),
    "fps": lambda: ts.farthest_point_sample(256, xyz),
    "gather": lambda: ts.gather_point(xyz, fps0),
    "ball resident": lambda: tg.query_ball_point(0.2, 32, xyz, new0),
    "ball stream": lambda: tg.query_ball_point(0.05, 16, big, big),
    "three_nn": lambda: ti.three_nn(xyz, new0),
    "linear_fwd simt": lambda: ffi.call("pn2_linear_fwd", 4096, 128, 128, p(A), 128, None, None, 0, p(W), None, p(Y), p(stats), None, 0, 0),
    "linear_fwd tc": lambda: ffi.call("pn2_linear_fwd", 4096, 128, 128, p(A), 128, None, None, 0, p(W), None, p(Y), p(stats), p(ws), ws.numel() * 4, 1),
    "linear_dgrad tc": lambda: ffi.call("pn2_linear_dgrad", 4096, 128, 128, p(A), p(W), p(dX), 128, p(ws), ws.numel() * 4, 1),
    "wgrad simt": lambda: ffi.call("pn2_linear_wgrad", 4096, 128, 128, p(A), 128, None, None, 0, p(Y), p(dW), None, 0),
    "wgrad tc": lambda: ffi.call("pn2_linear_wgrad", 131072, 64, 64, p(A2), 64, None, None, 0, p(dY2), p(dW2), None, 1),
    "wgrad rt": lambda: ffi.call("pn2_linear_wgrad", 8192, 32, 32, p(A3), 32, None, None, 0, p(dY3), p(dW3), None, 0),
    "bn_bwd_reduce": lambda: ffi.call("pn2_bn_bwd_reduce", 4096, 128, p(A), 128, p(Y), p(sc), p(sh), p(saved), 1, p(red)),
    "gather_grad (memset)": lambda: ts.gather_point_grad(xyz, fps0, new0),
}
fps0 = ts.farthest_point_sample(256, xyz); new0 = ts.gather_point(xyz, fps0)
big = torch.rand(64, 4096, 3, device=dev)
for name, fn in ops.items():
    fn(); torch.cuda.synchronize()       # warm (module load, attributes)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("%-24s capture OK" % name)
    except Exception as e:
        print("%-24s FAILED: %s" % (name, str(e).split("\n")[0][:150]))
        torch.cuda.synchronize()
