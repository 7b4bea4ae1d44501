"""A/B of an FPS variant selected by an environment switch (read once per process): times the SA1 call of config 2
(16 x 8192 -> 1024) and writes the indices, so that two runs can be compared bit for bit.
    python scripts/ab_fps.py out.npy"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200.tf_ops import tf_sampling as ts
rs = np.random.RandomState(100)
x = torch.as_tensor((rs.random_sample((16, 8192, 3)) * [10, 10, 5] - [5, 5, 0]).astype(np.float32)).cuda()
g = torch.as_tensor(rs.randint(0, 6, (4, 8192, 3)).astype(np.float32)).cuda()   # tie lattice
flush = torch.empty(64 * 1024 * 1024, device="cuda")
idx = ts.farthest_point_sample(1024, x); torch.cuda.synchronize()
t = []
for _ in range(7):
    flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx = ts.farthest_point_sample(1024, x); e1.record(); torch.cuda.synchronize()
    t.append(e0.elapsed_time(e1))
np.save(sys.argv[1], np.concatenate([idx.cpu().numpy().ravel(), ts.farthest_point_sample(700, g).cpu().numpy().ravel()]))
print("fps 16x8192->1024: %.4f ms (median of 7), %.3f us per round, PN2_FPS_T=%s" % (float(np.median(t)), float(np.median(t)) * 1e3 / 1023, os.environ.get("PN2_FPS_T")))
