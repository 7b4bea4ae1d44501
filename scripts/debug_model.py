"""Diagnostic (not a test): per-module forward error of the full model against the fp64 oracle,
isolated (each module fed the oracle's own inputs) and chained (our outputs fed forward)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pn2_b200  # noqa: E402
from pn2_b200.util import pointnet_util as pu, tf_util  # noqa: E402
from oracle import layers_ref as lr  # noqa: E402
from test_layers_gpu import HP_SMALL, load_params, randomize_bn  # noqa: E402


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def err(a, b):
    a = a.detach().cpu().numpy().astype(np.float64)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else b
    return float(np.abs(a - b).max()), float(np.abs(b).max())


def poison():
    """fill the caching allocator's free blocks with NaN so that reads of never-written padding show"""
    if os.environ.get("PN2_POISON") == "1":
        t = [torch.full((1 << 24,), float("nan"), device="cuda") for _ in range(8)]
        torch.cuda.synchronize()
        del t


def main(hp, b, n, scale):
    rs = np.random.RandomState(100)
    pc = np.concatenate([rs.random_sample((b, n, 3)) * np.asarray(scale), rs.random_sample((b, n, 3))],
                        -1).astype(np.float32)
    params = lr.init_model_params(hp, 9, seed=1)
    randomize_bn(params, rs)
    store = tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
    load_params(store, params)
    ctx = lr.Ctx(params, is_training=True, bn_decay=0.5)
    # oracle, module by module
    xyz = {0: np.ascontiguousarray(pc[:, :, :3])}
    pts = {0: torch.tensor(pc[:, :, 3:6], dtype=torch.float64)}
    for l in (1, 2, 3, 4):
        xyz[l], pts[l], _ = lr.sa_module(ctx, xyz[l - 1], pts[l - 1], hp["l%d_npoint" % l],
                                         hp["l%d_radius" % l], hp["l%d_nsample" % l], lr.SA_MLPS[l],
                                         "layer%d" % l)
    ups = {0: pts[4]}
    for l, (lo, hi) in zip((1, 2, 3, 4), ((3, 4), (2, 3), (1, 2), (0, 1))):
        ups[l] = lr.fp_module(ctx, xyz[lo], xyz[hi], pts[lo], ups[l - 1], lr.FP_MLPS[l], "fa_layer%d" % l)
    # ours: isolated and chained
    my_xyz, my_pts = {0: cu(xyz[0])}, {0: cu(pts[0].numpy())}
    for l in (1, 2, 3, 4):
        args = (hp["l%d_npoint" % l], hp["l%d_radius" % l], hp["l%d_nsample" % l], list(lr.SA_MLPS[l]),
                None, False, True, 0.5, "layer%d" % l)
        poison()
        _, iso, _ = pu.pointnet_sa_module(cu(xyz[l - 1]), cu(pts[l - 1].detach().numpy()), *args)
        my_xyz[l], my_pts[l], _ = pu.pointnet_sa_module(my_xyz[l - 1], my_pts[l - 1], *args)
        print("SA%d  isolated %.3g  chained %.3g  (|ref|max %.3g)  M=%d" % (
            l, err(iso, pts[l])[0], err(my_pts[l], pts[l])[0], err(iso, pts[l])[1],
            b * hp["l%d_npoint" % l] * hp["l%d_nsample" % l]))
        for i in range(3):
            key = "layer%d/conv%d" % (l, i)
            a = ctx.acts[key].detach().numpy()
            flat = a.reshape(-1, a.shape[-1])
            print("     %s post-act mean/std range: mean %.3g..%.3g std %.3g..%.3g" % (
                key, flat.mean(0).min(), flat.mean(0).max(), flat.std(0).min(), flat.std(0).max()))
    my_up = my_pts[4]
    for l, (lo, hi) in zip((1, 2, 3, 4), ((3, 4), (2, 3), (1, 2), (0, 1))):
        poison()
        iso = pu.pointnet_fp_module(cu(xyz[lo]), cu(xyz[hi]), cu(pts[lo].detach().numpy()),
                                    cu(ups[l - 1].detach().numpy()), list(lr.FP_MLPS[l]), True, 0.5,
                                    "fa_layer%d" % l)
        my_up = pu.pointnet_fp_module(my_xyz[lo], my_xyz[hi], my_pts[lo], my_up, list(lr.FP_MLPS[l]),
                                      True, 0.5, "fa_layer%d" % l)
        print("FP%d  isolated %.3g  chained %.3g  (|ref|max %.3g)  M=%d" % (
            l, err(iso, ups[l])[0], err(my_up, ups[l])[0], err(iso, ups[l])[1], b * xyz[lo].shape[1]))


if __name__ == "__main__":
    print("== small model")
    main(HP_SMALL, 2, 1024, (1.0, 1.0, 1.0))
    print("== semantic.json shape")
    hp = {"use_color": 1, "l1_npoint": 1024, "l1_radius": 0.5, "l1_nsample": 32, "l2_npoint": 256,
          "l2_radius": 1.0, "l2_nsample": 32, "l3_npoint": 64, "l3_radius": 2.0, "l3_nsample": 32,
          "l4_npoint": 16, "l4_radius": 4.0, "l4_nsample": 32}
    main(hp, 2, 8192, (10.0, 10.0, 5.0))
