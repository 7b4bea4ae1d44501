"""Diagnostic: SA+FP composite gradients vs fp64 oracle, vs an fp32 CPU oracle (conditioning),
and run-to-run determinism of our own backward."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pn2_b200
from pn2_b200.util import pointnet_util as pu, tf_util
from oracle import layers_ref as lr


def build_params(rs, rand_bn):
    params = {}
    k = 6
    for i, n in enumerate([32, 32, 64]):
        lr.init_conv(params, rs, "layer1/conv%d" % i, k, n); k = n
    k = 64 + 3
    for i, n in enumerate([64, 32]):
        lr.init_conv(params, rs, "fa/conv_%d" % i, k, n); k = n
    if rand_bn:
        for kk in list(params):
            if kk.endswith("/bn/gamma"): params[kk] = rs.uniform(0.5, 1.5, params[kk].shape).astype(np.float32)
            if kk.endswith("/bn/beta") or kk.endswith("/biases"): params[kk] = rs.uniform(-0.3, 0.3, params[kk].shape).astype(np.float32)
    return params


def oracle_grads(params, xyz, pts, gmat, dt):
    lr.set_dtype(dt)
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.5)
    pr = torch.tensor(pts, dtype=dt, requires_grad=True)
    e_xyz, e_feat, _ = lr.sa_module(ctx, xyz, pr, 256, 0.2, 32, [32, 32, 64], "layer1")
    e_out = lr.fp_module(ctx, xyz, e_xyz, pr, e_feat, [64, 32], "fa")
    (e_out * torch.tensor(gmat, dtype=dt)).sum().backward()
    lr.set_dtype(torch.float64)
    return ctx.grads(), e_out.detach().double().numpy()


def our_grads(params, xyz, pts, gmat):
    store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
    sd = {kk: (v.reshape((1, 1) + v.shape) if kk.endswith("weights") else v) for kk, v in params.items()}
    store.load_state_dict(sd)
    x = torch.as_tensor(xyz).cuda()
    pt = torch.as_tensor(pts).cuda().requires_grad_(True)
    l1_xyz, l1_feat, _ = pu.pointnet_sa_module(x, pt, 256, 0.2, 32, [32, 32, 64], None, False, True, 0.5, "layer1")
    out = pu.pointnet_fp_module(x, l1_xyz, pt, l1_feat, [64, 32], True, 0.5, "fa")
    (out * torch.as_tensor(gmat).cuda()).sum().backward()
    return {k: v.grad.cpu().numpy().astype(np.float64).reshape(-1) for k, v in store.vars.items() if v.grad is not None}, out.detach().cpu().numpy()


for rand_bn in (False, True):
    for loss in ("ones", "random"):
        rs = np.random.RandomState(100)
        xyz = rs.random_sample((2, 1024, 3)).astype(np.float32)
        pts = rs.random_sample((2, 1024, 3)).astype(np.float32)
        params = build_params(rs, rand_bn)
        gmat = np.ones((2, 1024, 32), np.float32) if loss == "ones" else rs.normal(size=(2, 1024, 32)).astype(np.float32)
        g64, o64 = oracle_grads(params, xyz, pts, gmat, torch.float64)
        g32, o32 = oracle_grads(params, xyz, pts, gmat, torch.float32)
        ga, oa = our_grads(params, xyz, pts, gmat)
        gb, ob = our_grads(params, xyz, pts, gmat)
        print("=== rand_bn=%s loss=%s  fwd err ours %.3g  fp32cpu %.3g" % (rand_bn, loss, np.abs(oa - o64).max(), np.abs(o32 - o64).max()))
        for k in g64:
            e = g64[k].reshape(-1)
            if k.endswith("biases"): continue
            print("  %-28s |g|max %9.4g  ours-vs-64 %9.3g  fp32cpu-vs-64 %9.3g  ours run-to-run %9.3g" % (
                k, np.abs(e).max(), np.abs(ga[k] - e).max(), np.abs(g32[k].reshape(-1) - e).max(), np.abs(ga[k] - gb[k]).max()))
