"""GPU: whole layers through one native call (pn2_sa_forward/backward, pn2_fp_forward/backward, SURVEY.md
section 8b) against the fp64 oracle AND against the Python-orchestrated layers on the same variables."""
import numpy as np
import pytest

from _util import to_cuda
from test_layers_gpu import ATOL, as_numpy, check_input_grad, check_param_grads, load_params, randomize_bn, \
    recorded_decisions

pytestmark = pytest.mark.gpu


@pytest.fixture()
def env(cuda):
    import pn2_b200  # noqa: F401
    from pn2_b200 import fused
    from pn2_b200.util import pointnet_util, tf_util
    from oracle import layers_ref as lr
    store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
    return fused, pointnet_util, tf_util, lr, store


@pytest.mark.parametrize("b,n,c,npoint,radius,ns,mlp", [(2, 1024, 3, 256, 0.2, 32, [32, 32, 64]),
                                                         (2, 512, 64, 128, 0.3, 32, [64, 64, 128]),
                                                         (3, 300, 0, 50, 0.4, 16, [16, 24])])
def test_sa_layer_single_call(env, b, n, c, npoint, radius, ns, mlp):
    fused, pu, tf_util, lr, store = env
    import torch
    rs = np.random.RandomState(n + c)
    xyz = rs.random_sample((b, n, 3)).astype(np.float32)
    pts = rs.random_sample((b, n, c)).astype(np.float32) if c else None
    params = {}
    k = 3 + c
    for i, w in enumerate(mlp):
        lr.init_conv(params, rs, "layer1/conv%d" % i, k, w)
        k = w
    randomize_bn(params, rs)
    load_params(store, params)
    # the Python-orchestrated layer first (records the decisions; moving statistics restored afterwards)
    keep = {k2: v.data.clone() for k2, v in store.vars.items() if not v.trainable}
    pt = None if pts is None else to_cuda(pts)
    with recorded_decisions(tf_util) as dec:
        p_xyz, p_out, p_idx = pu.pointnet_sa_module(to_cuda(xyz), pt, npoint, radius, ns, mlp, None, False, True,
                                                    0.7, "layer1")
    for k2, v in keep.items():
        store.vars[k2].data.copy_(v)
    sa = fused.SetAbstraction("layer1", c, npoint, radius, ns, mlp)
    new_xyz, out, idx = sa.forward(to_cuda(xyz), pt, is_training=True, bn_decay=0.7)
    assert bool((idx == p_idx).all()) and bool((new_xyz == p_xyz).all())
    np.testing.assert_allclose(out.cpu().numpy(), p_out.detach().cpu().numpy(), atol=2e-6)
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.7, decisions=as_numpy(dec))
    ref_pts = None if pts is None else torch.tensor(pts, dtype=torch.float64, requires_grad=True)
    e_xyz, e_out, e_idx = lr.sa_module(ctx, xyz, ref_pts, npoint, radius, ns, mlp, "layer1")
    np.testing.assert_array_equal(idx.cpu().numpy(), e_idx)
    np.testing.assert_allclose(out.cpu().numpy(), e_out.detach().numpy(), atol=ATOL)
    for k2, v in ctx.new_moving.items():
        np.testing.assert_allclose(store.vars[k2].data.cpu().numpy(), v, atol=ATOL, err_msg=k2)
    g = rs.normal(size=tuple(out.shape)).astype(np.float32)
    e_out.backward(torch.tensor(g, dtype=torch.float64))
    store.zero_grad()
    d_pts = sa.backward(to_cuda(g))
    check_param_grads(store, ctx)
    if c:
        check_input_grad(d_pts, ref_pts.grad.numpy(), ctx)


@pytest.mark.parametrize("b,n1,n2,c1,c2,mlp", [(2, 512, 64, 5, 24, [64, 32]), (2, 1024, 256, 64, 128, [128, 128]),
                                                (3, 200, 30, 0, 16, [16])])
def test_fp_layer_single_call(env, b, n1, n2, c1, c2, mlp):
    fused, pu, tf_util, lr, store = env
    import torch
    rs = np.random.RandomState(n1 + c1)
    xyz1 = rs.random_sample((b, n1, 3)).astype(np.float32)
    xyz2 = np.ascontiguousarray(xyz1[:, :n2])
    p1 = rs.random_sample((b, n1, c1)).astype(np.float32) if c1 else None
    p2 = rs.random_sample((b, n2, c2)).astype(np.float32)
    params = {}
    k = c1 + c2
    for i, w in enumerate(mlp):
        lr.init_conv(params, rs, "fa/conv_%d" % i, k, w)
        k = w
    randomize_bn(params, rs)
    load_params(store, params)
    fp = fused.FeaturePropagation("fa", c1, c2, mlp)
    out = fp.forward(to_cuda(xyz1), to_cuda(xyz2), None if p1 is None else to_cuda(p1), to_cuda(p2), True, 0.9)
    # decisions: the chain's masks are a function of its pre-activations, which the Python path reproduces
    keep = {k2: v.data.clone() for k2, v in store.vars.items() if not v.trainable}
    with recorded_decisions(tf_util) as dec:
        p_out = pu.pointnet_fp_module(to_cuda(xyz1), to_cuda(xyz2), None if p1 is None else to_cuda(p1), to_cuda(p2),
                                      mlp, True, 0.9, "fa")
    for k2, v in keep.items():
        store.vars[k2].data.copy_(v)
    np.testing.assert_allclose(out.cpu().numpy(), p_out.detach().cpu().numpy(), atol=2e-6)
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.9, decisions=as_numpy(dec))
    r1 = None if p1 is None else torch.tensor(p1, dtype=torch.float64, requires_grad=True)
    r2 = torch.tensor(p2, dtype=torch.float64, requires_grad=True)
    e_out = lr.fp_module(ctx, xyz1, xyz2, r1, r2, mlp, "fa")
    np.testing.assert_allclose(out.cpu().numpy(), e_out.detach().numpy(), atol=ATOL)
    g = rs.normal(size=tuple(out.shape)).astype(np.float32)
    e_out.backward(torch.tensor(g, dtype=torch.float64))
    store.zero_grad()
    d1, d2 = fp.backward(to_cuda(g))
    check_param_grads(store, ctx)
    check_input_grad(d2, r2.grad.numpy(), ctx)
    if c1:
        check_input_grad(d1, r1.grad.numpy(), ctx)


def test_fused_layer_validation(env):
    fused, _, _, _, _ = env
    import ctypes
    import pn2_b200
    L = pn2_b200._ffi.lib()
    cfg = fused.SaConfig(2, 100, 3, 10, 8, 9, 1, 0.2, 1e-3, 0.9)          # 9 layers > PN2_MAX_LAYERS
    assert L.pn2_sa_workspace_bytes(ctypes.byref(cfg), (fused.ConvLayer * 9)()) == -1
    cfg = fused.SaConfig(2, 100, 3, 10, 8, 1, 1, -1.0, 1e-3, 0.9)         # radius <= 0
    lay = (fused.ConvLayer * 1)()
    lay[0].K, lay[0].N = 6, 8
    assert L.pn2_sa_workspace_bytes(ctypes.byref(cfg), lay) == -1
    cfg = fused.SaConfig(2, 100, 3, 10, 8, 1, 1, 0.2, 1e-3, 0.9)
    lay[0].K = 7                                                           # first layer must take 3 + c channels
    assert L.pn2_sa_workspace_bytes(ctypes.byref(cfg), lay) == -1
    lay[0].K = 6
    assert L.pn2_sa_workspace_bytes(ctypes.byref(cfg), lay) > 0
