"""Diagnostic: chained SA/FP modules, then conv0 of FP4 / SA3 called directly in tensor-core and fp32
modes on the very same padded buffer; prints where the two differ."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pn2_b200
from pn2_b200.util import pointnet_util as pu, tf_util
from pn2_b200 import _ffi as ffi
from oracle import layers_ref as lr
from test_layers_gpu import load_params, randomize_bn
p = ffi.ptr

hp = {"use_color": 1, "l1_npoint": 1024, "l1_radius": 0.5, "l1_nsample": 32, "l2_npoint": 256,
      "l2_radius": 1.0, "l2_nsample": 32, "l3_npoint": 64, "l3_radius": 2.0, "l3_nsample": 32,
      "l4_npoint": 16, "l4_radius": 4.0, "l4_nsample": 32}
b, n = 2, 8192
rs = np.random.RandomState(100)
pc = np.concatenate([rs.random_sample((b, n, 3)) * np.asarray((10., 10., 5.)), rs.random_sample((b, n, 3))], -1).astype(np.float32)
params = lr.init_model_params(hp, 9, seed=1); randomize_bn(params, rs)
store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0)); load_params(store, params)
if os.environ.get("PN2_POISON") == "1":
    t = [torch.full((1 << 24,), float("nan"), device="cuda") for _ in range(8)]; torch.cuda.synchronize(); del t
xyz0 = torch.as_tensor(pc[:, :, :3]).cuda().contiguous(); pts0 = torch.as_tensor(pc[:, :, 3:6]).cuda().contiguous()
xyz, pts = {0: xyz0}, {0: pts0}
for l in (1, 2):
    xyz[l], pts[l], _ = pu.pointnet_sa_module(xyz[l - 1], pts[l - 1], hp["l%d_npoint" % l], hp["l%d_radius" % l],
                                             hp["l%d_nsample" % l], list(lr.SA_MLPS[l]), None, False, True, 0.5, "layer%d" % l)
# SA3 conv0 by hand on the chained inputs
new_xyz, new_points, idx, _ = pu.sample_and_group(64, 2.0, 32, xyz[2], pts[2], False, True)
bb, m, ns, k = new_points.shape
x2d = new_points.reshape(bb * m * ns, k)
print("x2d shape", tuple(x2d.shape), "strides", x2d.stride(), "ptr%16", x2d.data_ptr() % 16, "requires_grad", x2d.requires_grad)
L = tf_util.make_layer("dbg/conv0", k, 128, True, tf_util.relu)
M, N = x2d.shape[0], 128
ws = torch.empty(int(ffi.lib().pn2_linear_workspace_bytes(k, N)) // 4 + 4, device="cuda")
def run(mode, x):
    Y = torch.empty(M, N, device="cuda")
    a_ptr, lda = ffi.ptr_rows(x.detach(), torch.float32)
    ffi.call("pn2_linear_fwd", M, k, N, a_ptr, lda, None, None, 0, p(L.w.data), p(L.b.data), p(Y), None, p(ws), ws.numel() * 4, mode)
    torch.cuda.synchronize()
    return Y
Y1 = run(1, x2d); Y0 = run(0, x2d); Y1c = run(1, x2d.detach().contiguous())
for rep in range(3):
    Yr = run(1, x2d)
    dr = (Yr - Y0).abs()
    rows_r = (dr > 1e-4).nonzero()[:, 0].unique()
    print("repeat %d: max %.4g bad rows %s" % (rep, dr.max().item(), rows_r[:12].tolist()))
# same data, padding forced to zero
base = x2d.detach()
full = torch.as_strided(base, (M, base.stride(0)), (base.stride(0), 1), base.storage_offset())
print("padding column finite?", torch.isfinite(full[:, k:]).all().item(), " any NaN in valid cols?", torch.isnan(full[:, :k]).any().item())
full2 = full.clone(); full2[:, k:] = 0
Yz = run(1, full2[:, :k])
print("tc with zeroed padding vs fp32: max %.4g" % (Yz - Y0).abs().max().item())
full3 = full.clone(); full3[:, k:] = float("nan")
Yn = run(1, full3[:, :k])
dn = (Yn - Y0).abs(); print("tc with NaN padding vs fp32: max %.4g  nan count %d" % (torch.nan_to_num(dn, nan=0.0).max().item(), torch.isnan(Yn).sum().item()))
d = (Y1 - Y0).abs()
print("tc vs fp32 on padded view: max", d.max().item(), " tc(contiguous copy) vs fp32:", (Y1c - Y0).abs().max().item())
bad = (d > 1e-4).nonzero()
print("bad elements", bad.shape[0], "of", d.numel())
if bad.shape[0]:
    rows = bad[:, 0].unique(); cols = bad[:, 1].unique()
    print("bad rows: n=%d min %d max %d first %s" % (rows.numel(), rows.min().item(), rows.max().item(), rows[:16].tolist()))
    print("bad cols: n=%d first %s" % (cols.numel(), cols[:16].tolist()))
    # which K chunk explains the difference?
    xa = x2d.detach().double(); W = L.w.data.reshape(k, N).double()
    r = rows[0].item()
    for kc in range((k + 31) // 32):
        contrib = xa[r, kc * 32:(kc + 1) * 32] @ W[kc * 32:(kc + 1) * 32]
        print("  row %d chunk %d contrib to col0 %.4f" % (r, kc, contrib[0].item()))
    print("  row %d: tc %.4f fp32 %.4f" % (r, Y1[r, 0].item(), Y0[r, 0].item()))
