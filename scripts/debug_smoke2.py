"""Diagnostic: per-channel dbeta of fa/conv_0 in the SA+FP composite with degenerate BN params."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pn2_b200
from pn2_b200.util import pointnet_util as pu, tf_util
from oracle import layers_ref as lr
from debug_smoke import build_params

rs = np.random.RandomState(100)
xyz = rs.random_sample((2, 1024, 3)).astype(np.float32)
pts = rs.random_sample((2, 1024, 3)).astype(np.float32)
params = build_params(rs, False)
gmat = rs.normal(size=(2, 1024, 32)).astype(np.float32)
ctx = lr.Ctx(params, is_training=True, bn_decay=0.5)
pr = torch.tensor(pts, dtype=torch.float64, requires_grad=True)
e_xyz, e_feat, _ = lr.sa_module(ctx, xyz, pr, 256, 0.2, 32, [32, 32, 64], "layer1")
e_out = lr.fp_module(ctx, xyz, e_xyz, pr, e_feat, [64, 32], "fa")
(e_out * torch.tensor(gmat, dtype=torch.float64)).sum().backward()
g64 = ctx.grads()

store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
store.load_state_dict({kk: (v.reshape((1, 1) + v.shape) if kk.endswith("weights") else v) for kk, v in params.items()})
x = torch.as_tensor(xyz).cuda()
pt = torch.as_tensor(pts).cuda().requires_grad_(True)
l1_xyz, l1_feat, _ = pu.pointnet_sa_module(x, pt, 256, 0.2, 32, [32, 32, 64], None, False, True, 0.5, "layer1")
l1_feat.retain_grad()
out = pu.pointnet_fp_module(x, l1_xyz, pt, l1_feat, [64, 32], True, 0.5, "fa")
(out * torch.as_tensor(gmat).cuda()).sum().backward()
print("l1_feat fwd err", float(np.abs(l1_feat.detach().cpu().numpy() - e_feat.detach().numpy()).max()))
print("exact zeros in l1_feat: ours %d oracle %d of %d" % ((l1_feat == 0).sum().item(), (e_feat == 0).sum().item(), l1_feat.numel()))
# oracle pre-BN activations of fa/conv_0
from oracle import oracle as orc
dist, idx = orc.three_nn(xyz, e_xyz)
w = lr.fp_weights(dist)
ef = e_feat.detach().numpy()
rows = np.stack([np.stack([ef[b][idx[b, :, t]] for t in range(3)], 1) for b in range(2)])  # b,n,3,c
interp = (rows * w[..., None]).sum(2)
x0 = np.concatenate([interp, pts.astype(np.float64)], -1).reshape(-1, 67)
y0 = x0 @ params["fa/conv_0/weights"].astype(np.float64)
d = np.abs(store.vars["fa/conv_0/bn/beta"].grad.cpu().numpy() - g64["fa/conv_0/bn/beta"])
print("x0 col std min %.3g ; constant x0 columns: %s" % (x0.std(0).min(), np.where(x0.std(0) < 1e-9)[0]))
for c in np.argsort(-d)[:6]:
    zz = (y0[:, c] - y0[:, c].mean()) / np.sqrt(y0[:, c].var() + 1e-3)
    print("ch %2d dbeta ours %.6g oracle %.6g | y mean %.3g std %.3g | #|z|<1e-6 %d #|z|<1e-4 %d #|z|<1e-2 %d" % (
        c, store.vars["fa/conv_0/bn/beta"].grad[c].item(), g64["fa/conv_0/bn/beta"][c], y0[:, c].mean(), y0[:, c].std(),
        (np.abs(zz) < 1e-6).sum(), (np.abs(zz) < 1e-4).sum(), (np.abs(zz) < 1e-2).sum()))
# gradient wrt l1_feat and pts
eg = e_feat.grad if e_feat.grad is not None else None
print("d l1_feat: ours-vs-oracle", None if eg is None else float(np.abs(l1_feat.grad.cpu().numpy() - eg.numpy()).max()))
print("d pts: ours-vs-oracle %.3g (|g|max %.3g)" % (float(np.abs(pt.grad.cpu().numpy() - pr.grad.numpy()).max()), float(pr.grad.abs().max())))
