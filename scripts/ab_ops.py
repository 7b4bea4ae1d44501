"""A/B timing of the round-1 "experimental" kernels against the default ones at config-2 sizes
(B=16 clouds): fp32-gated 3-NN vs the fp64 kernel, grid ball query vs the TMA brute-force kernels.
CUDA events on the launching stream, 256 MB L2 flush before every timed call, median of 7."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200._ffi import F32, I32, call, lib, ptr
from pn2_b200.tf_ops import tf_grouping as tg, tf_interpolate as ti, tf_sampling as ts

dev = "cuda"
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return float(np.median(t))

rs = np.random.RandomState(100)
B = 16
xyz0 = torch.as_tensor((rs.random_sample((B, 8192, 3)) * [10, 10, 5] - [5, 5, 0]).astype(np.float32)).to(dev)
levels = [xyz0]
for m in (1024, 256, 64, 16):
    levels.append(ts.gather_point(levels[-1], ts.farthest_point_sample(m, levels[-1])))
out = []
def nn_filtered(x1, x2):
    b, n, _ = x1.shape
    d = torch.empty((b, n, 3), dtype=F32, device=dev); i = torch.empty((b, n, 3), dtype=I32, device=dev)
    call("pn2_three_nn_filtered", b, n, x2.shape[1], ptr(x1, F32), ptr(x2, F32), ptr(d, F32), ptr(i, I32))
    return d, i
for lo in (0, 1, 2, 3):
    x1, x2 = levels[lo], levels[lo + 1]
    a = timed(lambda: ti.three_nn(x1, x2)); b = timed(lambda: nn_filtered(x1, x2))
    d0, i0 = ti.three_nn(x1, x2); d1, i1 = nn_filtered(x1, x2)
    out.append({"op": "three_nn", "n": x1.shape[1], "m": x2.shape[1], "default_ms": a, "filtered_ms": b,
                "identical": bool((i0 == i1).all() and (d0 == d1).all())})
    print(out[-1])
def grid(radius, ns, x1, x2, ws, nbytes):
    b, n, _ = x1.shape; m = x2.shape[1]
    idx = torch.empty((b, m, ns), dtype=I32, device=dev); cnt = torch.empty((b, m), dtype=I32, device=dev)
    call("pn2_query_ball_point_grid", b, n, m, float(radius), ns, ptr(x1, F32), ptr(x2, F32), ptr(idx, I32),
         ptr(cnt, I32), ptr(ws, F32), nbytes)
    return idx, cnt
for l, r in ((0, 0.5), (1, 1.0), (2, 2.0), (3, 4.0)):
    x1, x2 = levels[l], levels[l + 1]
    nbytes = int(lib().pn2_ball_grid_workspace_bytes(B, x1.shape[1]))
    ws = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.float32, device=dev)
    a = timed(lambda: tg.query_ball_point(r, 32, x1, x2)); b = timed(lambda: grid(r, 32, x1, x2, ws, nbytes))
    i0, c0 = tg.query_ball_point(r, 32, x1, x2); i1, c1 = grid(r, 32, x1, x2, ws, nbytes)
    out.append({"op": "query_ball_point", "n": x1.shape[1], "m": x2.shape[1], "radius": r, "default_ms": a,
                "grid_ms": b, "identical": bool((i0 == i1).all() and (c0 == c1).all())})
    print(out[-1])
# FPS at config-2 sizes (for the record)
for l, m in ((0, 1024), (1, 256), (2, 64), (3, 16)):
    x = levels[l]
    a = timed(lambda: ts.farthest_point_sample(m, x))
    out.append({"op": "fps", "n": x.shape[1], "m": m, "ms": a, "us_per_round": a * 1e3 / (m - 1)})
    print(out[-1])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_ops.json"), "w"), indent=1)
