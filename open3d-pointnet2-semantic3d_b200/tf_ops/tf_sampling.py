"""Sampling ops: same names / argument order as the reference's tf_ops/tf_sampling.py.

  prob_sample(inp, inpr)               tf_sampling.py:18-26  (no gradient, :29)
  farthest_point_sample(npoint, inp)   tf_sampling.py:61-69  (no gradient, :72)
  gather_point(inp, idx)               tf_sampling.py:38-46  (gradient :54-58)

Inputs and outputs are contiguous CUDA torch tensors (float32 / int32); the work is done by
the sm_100a kernels behind the C ABI (csrc/pn2_sampling.cu).  Shape errors raise ValueError
with the reference's InvalidArgument texts (tf_sampling.cpp:86-96, 121-132, 166-178).
"""
import torch

from .. import _ffi
from .._ffi import F32, I32, call, ptr


def _need(cond, msg):
    if not cond:
        raise ValueError(msg)


def prob_sample(inp, inpr):
    """inp (B,ncategory) float32 weights, inpr (B,npoints) float32 uniform numbers in [0,1) ->
    (B,npoints) int32: category drawn by inverting the cumulative weights.  The prefix sum follows
    the reference's fp32 addition order, so the result is bit-identical (csrc/pn2_sampling.cu)."""
    _need(inp.dim() == 2, "ProbSample expects (batch_size,num_choices) inp shape")
    _need(inpr.dim() == 2 and inpr.shape[0] == inp.shape[0],
          "ProbSample expects (batch_size,num_points) inpr shape")
    b, n = inp.shape
    m = inpr.shape[1]
    _need(n > 0, "ProbSample expects (batch_size,num_choices) inp shape")
    inp, inpr = inp.detach().contiguous(), inpr.detach().contiguous()
    out = torch.empty((b, m), dtype=I32, device=inp.device)
    temp = torch.empty((b, n), dtype=F32, device=inp.device)
    call("pn2_prob_sample", b, n, m, ptr(inp, F32), ptr(inpr, F32), ptr(temp, F32), ptr(out, I32))
    return out


def farthest_point_sample(npoint, inp):
    """inp (B,N,3) float32 -> (B,npoint) int32 indices; seed index 0, reference tie order."""
    _need(int(npoint) > 0, "FarthestPointSample expects positive npoint")
    _need(inp.dim() == 3 and inp.shape[2] == 3,
          "FarthestPointSample expects (batch_size,num_points,3) inp shape")
    b, n, _ = inp.shape
    inp = inp.detach().contiguous()
    out = torch.empty((b, int(npoint)), dtype=I32, device=inp.device)
    temp = None
    if n > 16384:  # streaming kernel keeps the running minimum in global memory
        temp = torch.empty((b, n), dtype=F32, device=inp.device)
    call("pn2_fps", b, n, int(npoint), ptr(inp, F32), ptr(temp, F32, allow_none=True),
         ptr(out, I32))
    return out


class _GatherPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, idx):
        b, n, _ = inp.shape
        m = idx.shape[1]
        inp_c = inp.contiguous()
        out = torch.empty((b, m, 3), dtype=F32, device=inp.device)
        call("pn2_gather_point", b, n, m, ptr(inp_c, F32), ptr(idx, I32), ptr(out, F32))
        ctx.save_for_backward(idx)
        ctx.dims = (b, n, m)
        return out

    @staticmethod
    def backward(ctx, out_g):
        (idx,) = ctx.saved_tensors
        b, n, m = ctx.dims
        out_g = out_g.contiguous()
        inp_g = torch.empty((b, n, 3), dtype=F32, device=out_g.device)
        call("pn2_gather_point_grad", b, n, m, ptr(out_g, F32), ptr(idx, I32), ptr(inp_g, F32))
        return inp_g, None


def gather_point(inp, idx):
    """inp (B,N,3) float32, idx (B,m) int32 -> (B,m,3); differentiable w.r.t. inp."""
    _need(inp.dim() == 3 and inp.shape[2] == 3,
          "GatherPoint expects (batch_size,num_points,3) inp shape")
    _need(idx.dim() == 2 and idx.shape[0] == inp.shape[0],
          "GatherPoint expects (batch_size,num_result) idx shape")
    return _GatherPoint.apply(inp, idx.contiguous())


def gather_point_grad(inp, idx, out_g):
    """Explicit gradient op (reference: sampling_module.gather_point_grad)."""
    b, n, _ = inp.shape
    m = idx.shape[1]
    _need(out_g.dim() == 3 and out_g.shape[0] == b and out_g.shape[1] == m and out_g.shape[2] == 3,
          "GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape")
    inp_g = torch.empty((b, n, 3), dtype=F32, device=inp.device)
    og, ii = out_g.contiguous(), idx.contiguous()
    call("pn2_gather_point_grad", b, n, m, ptr(og, F32), ptr(ii, I32), ptr(inp_g, F32))
    return inp_g
