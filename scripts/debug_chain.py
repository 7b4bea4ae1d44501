"""Diagnostic: a dense 2-layer BN-ReLU chain (the FP-module MLP) with degenerate BN parameters,
per-channel comparison of dbeta/dgamma against torch fp64 on the SAME input."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200.util import tf_util

for seed in (0, 1, 2):
    rs = np.random.RandomState(seed)
    M, K0, N0, N1 = 2048, 67, 64, 32
    x = rs.normal(size=(M, K0)).astype(np.float32)
    x[:, :64] = np.abs(x[:, :64])
    if seed == 2:   # a quarter of the rows duplicated (as padded groups do)
        x[M // 2:] = x[: M // 2]
    W0 = (rs.uniform(-1, 1, (K0, N0)) * np.sqrt(6 / (K0 + N0))).astype(np.float32)
    W1 = (rs.uniform(-1, 1, (N0, N1)) * np.sqrt(6 / (N0 + N1))).astype(np.float32)
    g = rs.normal(size=(M, N1)).astype(np.float32)
    dt = torch.float64
    xt = torch.tensor(x, dtype=dt); w0 = torch.tensor(W0, dtype=dt, requires_grad=True); w1 = torch.tensor(W1, dtype=dt, requires_grad=True)
    b0 = torch.zeros(N0, dtype=dt, requires_grad=True); g0 = torch.ones(N0, dtype=dt, requires_grad=True)
    b1 = torch.zeros(N1, dtype=dt, requires_grad=True); g1 = torch.ones(N1, dtype=dt, requires_grad=True)
    def bn(y, ga, be):
        m = y.mean(0); v = y.var(0, unbiased=False); return (y - m) * torch.rsqrt(v + 1e-3) * ga + be
    y0 = xt @ w0; z0 = torch.relu(bn(y0, g0, b0)); z1 = torch.relu(bn(z0 @ w1, g1, b1))
    (z1 * torch.tensor(g, dtype=dt)).sum().backward()

    store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
    store.load_state_dict({"c0/weights": W0.reshape(1, 1, K0, N0), "c1/weights": W1.reshape(1, 1, N0, N1)})
    L0 = tf_util.make_layer("c0", K0, N0, True, tf_util.relu)
    L1 = tf_util.make_layer("c1", N0, N1, True, tf_util.relu)
    xin = torch.as_tensor(x).cuda().requires_grad_(True)
    out = tf_util.mlp_chain(xin, [L0, L1], True, 0.5)
    (out * torch.as_tensor(g).cuda()).sum().backward()
    print("seed", seed, "fwd err", float(np.abs(out.detach().cpu().numpy() - z1.detach().numpy()).max()))
    for name, ours, ref in [("c0/beta", L0.beta.grad, b0.grad), ("c0/gamma", L0.gamma.grad, g0.grad),
                            ("c1/beta", L1.beta.grad, b1.grad), ("c0/w", L0.w.grad, w0.grad),
                            ("c1/w", L1.w.grad, w1.grad), ("dx", xin.grad, None)]:
        if ref is None:
            continue
        d = np.abs(ours.cpu().numpy().reshape(ref.shape).astype(np.float64) - ref.numpy())
        print("   %-9s max|g| %8.3g  max diff %9.3g  at %s" % (name, float(ref.abs().max()), d.max(), np.unravel_index(d.argmax(), d.shape)))
    d = np.abs(L0.beta.grad.cpu().numpy() - b0.grad.numpy())
    worst = np.argsort(-d)[:4]
    yy = y0.detach().numpy()
    for c in worst:
        zz = (yy[:, c] - yy[:, c].mean()) / np.sqrt(yy[:, c].var() + 1e-3)
        print("   ch %2d dbeta diff %.3g  y mean %.3g std %.3g  #|z|<1e-5: %d  #|z|<1e-3: %d" % (
            c, d[c], yy[:, c].mean(), yy[:, c].std(), (np.abs(zz) < 1e-5).sum(), (np.abs(zz) < 1e-3).sum()))
