"""Generates the committed fixtures under tests/golden/.

The reference's TensorFlow ops cannot be imported here (TensorFlow and Open3D are absent), so
the only reference-produced numbers available are the ones its own test prints
(tf_ops/test_interpolate.py:30-35); they are embedded in tests/test_oracle_golden.py.  This
script freezes oracle outputs on the same seed-100 stream so that later changes to the oracle
(or a different numpy) are detected:
  three_nn_seed100.npz   first 256 queries of batch 0 of the reference's golden input
  fps_ball_seed100.npz   FPS(256) + ball query(0.2, 32) on BASELINE.json config 1
  prob_vote_seed100.npz  prob_sample on the reference smoke test's triangle areas
                         (tf_ops/test_tf_ops.py:96-110) + CDF of a 9000-entry weight row;
                         interpolate_label_with_color (knn 3) on a 700/400-point pair
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

np.random.seed(100)
target = np.random.random((64, 8192, 3)).astype("float32")
reference = np.random.random((64, 1024, 3)).astype("float32")
dist, idx = orc.three_nn(target[:1, :256], reference[:1])
np.savez_compressed(os.path.join(HERE, "three_nn_seed100.npz"), dist=dist, idx=idx)

rs = np.random.RandomState(100)
xyz = rs.random_sample((2, 1024, 3)).astype(np.float32)
fps = orc.farthest_point_sample(256, xyz)
new_xyz = orc.gather_point(xyz, fps)
bidx, bcnt = orc.query_ball_point(0.2, 32, xyz, new_xyz)
np.savez_compressed(os.path.join(HERE, "fps_ball_seed100.npz"), fps=fps, idx=bidx, cnt=bcnt)
np.random.seed(100)
tri = np.random.rand(1, 5, 3, 3).astype("float32")
ta, tb, tc = tri[:, :, 0], tri[:, :, 1], tri[:, :, 2]
areas = np.sqrt((np.cross(tb - ta, tc - ta) ** 2).sum(2) + 1e-9).astype(np.float32)
r = np.random.rand(1, 8192).astype(np.float32)
rs = np.random.RandomState(100)
w = rs.random_sample((1, 9000)).astype(np.float32)
sp = rs.random_sample((700, 3)).astype(np.float32)
sl = rs.randint(0, 9, 700).astype(np.int32)
dp = rs.random_sample((400, 3)).astype(np.float32)
vl, vc = orc.interpolate_label_with_color(sp, sl, dp, 3)
np.savez_compressed(os.path.join(HERE, "prob_vote_seed100.npz"), ids=orc.prob_sample(areas, r),
                    cdf=orc.cumsum(w), vote_labels=vl, vote_colors=vc)
print("wrote fixtures")
