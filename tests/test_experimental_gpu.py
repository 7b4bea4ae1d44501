"""GPU tests of kernels that exist but are NOT on any default path yet: they were written after the
round's GPU budget was spent, so nothing here has run on a B200.  Skipped unless PN2_EXPERIMENTAL=1
(tests/conftest.py); the first GPU call of the next round runs them, and a kernel is switched on by
default only after its test is green and its timing is recorded under profiles/."""
import numpy as np
import pytest

from _util import rng_cloud, to_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.experimental]


@pytest.fixture(scope="module")
def ops(cuda):
    import pn2_b200  # noqa: F401
    from pn2_b200.tf_ops import tf_grouping, tf_interpolate, tf_sampling
    from oracle import oracle as orc
    return tf_sampling, tf_grouping, tf_interpolate, orc


def _three_nn_filtered(x1, x2):
    import torch
    from pn2_b200._ffi import F32, I32, call, ptr
    b, n, _ = x1.shape
    m = x2.shape[1]
    dist = torch.empty((b, n, 3), dtype=F32, device=x1.device)
    idx = torch.empty((b, n, 3), dtype=I32, device=x1.device)
    call("pn2_three_nn_filtered", b, n, m, ptr(x1, F32), ptr(x2, F32), ptr(dist, F32), ptr(idx, I32))
    return dist, idx


def _nn_cases():
    rs = np.random.RandomState(0)
    u = lambda *s: rs.random_sample(s)  # noqa: E731
    return {
        "uniform": (u(2, 3000, 3), u(2, 2500, 3)),
        "cfg2_fp4": (u(16, 8192, 3) * [10, 10, 5], u(16, 1024, 3) * [10, 10, 5]),
        "offset_1e3": (1000 + 1e-3 * u(1, 700, 3), 1000 + 1e-3 * u(1, 4100, 3)),
        "offset_1e5": (1e5 + u(1, 700, 3), 1e5 + u(1, 2049, 3)),
        "lattice": (rs.randint(0, 3, (2, 500, 3)), rs.randint(0, 3, (2, 900, 3))),
        "tiny": (1e-22 * u(1, 300, 3), 1e-22 * u(1, 700, 3)),
        "duplicates": (np.repeat(u(1, 100, 3), 4, 1), np.repeat(u(1, 300, 3), 5, 1)),
        "huge": (1e18 * u(1, 300, 3), 1e18 * u(1, 800, 3)),
        "anisotropic": (u(1, 500, 3) * [1e-3, 1, 1e3], u(1, 3000, 3) * [1e-3, 1, 1e3]),
        "three_known": (u(2, 100, 3), u(2, 3, 3)),
    }


@pytest.mark.parametrize("name", sorted(_nn_cases()))
def test_three_nn_filtered_is_bit_identical(ops, name):
    """The fp32 gate may only skip pairs that cannot enter the top 3: indices and distances must equal
    the oracle's bit for bit, also under cancellation (large offsets), underflow (1e-22), overflow of
    the fp32 distance (1e18) and on tie lattices."""
    _, _, _, orc = ops
    x1, x2 = (np.ascontiguousarray(a, dtype=np.float32) for a in _nn_cases()[name])
    dist, idx = _three_nn_filtered(to_cuda(x1), to_cuda(x2))
    ed, ei = orc.three_nn(x1, x2, threads=8)
    np.testing.assert_array_equal(idx.cpu().numpy(), ei)
    np.testing.assert_array_equal(dist.cpu().numpy().view(np.uint32), ed.view(np.uint32))


def test_three_nn_filtered_matches_default_kernel_full_size(ops):
    _, _, ti, _ = ops
    x1, x2 = to_cuda(rng_cloud(1, 16, 8192)), to_cuda(rng_cloud(2, 16, 1024))
    d0, i0 = ti.three_nn(x1, x2)
    d1, i1 = _three_nn_filtered(x1, x2)
    assert bool((i0 == i1).all()) and bool((d0 == d1).all())


# --------------------------------------------------------------- cluster FPS with the mbarrier handshake
def _fps_cluster_mb(x, m):
    import torch
    from pn2_b200._ffi import F32, I32, call, ptr
    b, n, _ = x.shape
    out = torch.empty((b, m), dtype=I32, device=x.device)
    call("pn2_fps_cluster_mb", b, n, m, ptr(x, F32), ptr(out, I32))
    torch.cuda.synchronize()  # a wrong handshake shows up as a hang: run this file under `timeout`
    return out


@pytest.mark.parametrize("b,n,m", [(1, 12000, 64), (1, 20000, 32), (2, 16389, 100), (3, 40000, 150),
                                   (1, 65536, 300), (1, 262144, 200), (16, 16384, 64), (2, 100, 120)])
def test_fps_cluster_mb_matches_oracle(ops, b, n, m):
    _, _, _, orc = ops
    x = rng_cloud(300 + n, b, n)
    got = _fps_cluster_mb(to_cuda(x), m).cpu().numpy()
    np.testing.assert_array_equal(got, orc.farthest_point_sample(m, x, threads=8))


def test_fps_cluster_mb_tie_order_and_long_run(ops):
    """Tie lattice across CTAs, and many rounds (phase parity of the two mbarriers flips 4000 times)."""
    ts, _, _, orc = ops
    rs = np.random.RandomState(9)
    x = rs.randint(0, 6, (2, 20000, 3)).astype(np.float32)
    np.testing.assert_array_equal(_fps_cluster_mb(to_cuda(x), 300).cpu().numpy(),
                                  orc.farthest_point_sample(300, x, threads=8))
    y = to_cuda(rng_cloud(5, 1, 32768))
    np.testing.assert_array_equal(_fps_cluster_mb(y, 8000).cpu().numpy(),
                                  ts.farthest_point_sample(8000, y).cpu().numpy())


# ------------------------------------------------------------------ ball query over the hashed grid
def _ball_grid(radius, ns, x1, x2):
    import torch
    from pn2_b200._ffi import F32, I32, call, lib, ptr
    b, n, _ = x1.shape
    m = x2.shape[1]
    nbytes = int(lib().pn2_ball_grid_workspace_bytes(b, n))
    ws = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.float32, device=x1.device)
    idx = torch.empty((b, m, ns), dtype=I32, device=x1.device)
    cnt = torch.empty((b, m), dtype=I32, device=x1.device)
    call("pn2_query_ball_point_grid", b, n, m, float(radius), int(ns), ptr(x1, F32), ptr(x2, F32),
         ptr(idx, I32), ptr(cnt, I32), ptr(ws, F32), nbytes)
    torch.cuda.synchronize()
    return idx, cnt


def _grid_cases():
    rs = np.random.RandomState(1)

    def cloud(b, n, scale=(1, 1, 1), shift=(0, 0, 0)):
        return (rs.random_sample((b, n, 3)) * scale + shift).astype(np.float32)

    def queries(x, m):
        q = np.stack([xb[rs.choice(x.shape[1], m, replace=False)] for xb in x]).copy()
        q[:, :5] += np.float32(0.013)  # a few queries that are not data points
        return q

    cases = {}
    x = cloud(16, 8192, (10, 10, 5), (-5, -5, 0))
    cases["cfg2_sa1"] = (0.5, 32, x, queries(x, 1024))
    x = cloud(2, 3000)
    cases["unit_r0.2"] = (0.2, 32, x, queries(x, 300))
    cases["dense_over_cap"] = (0.6, 16, x, queries(x, 100))      # > 512 hits per ball: brute-force fallback
    x = cloud(2, 1000, (10, 10, 5), (-5, -5, 0))
    cases["radius_4"] = (4.0, 32, x, queries(x, 64))
    x = cloud(1, 3000, (1, 1, 1), (1000, -2000, 50))
    cases["offset"] = (0.2, 32, x, queries(x, 200))
    x = cloud(1, 2000)
    cases["tiny_radius"] = (1e-3, 8, x, queries(x, 200))
    x = rs.randint(0, 8, (2, 3000, 3)).astype(np.float32)
    cases["lattice_r1"] = (1.0, 32, x, queries(x, 200))            # distance exactly == radius: not a hit
    x = cloud(1, 16384)
    cases["cfg5_16k"] = (float((3.0 * 2 * 64 / (4.0 * np.pi * 16384)) ** (1.0 / 3.0)), 64, x, queries(x, 4096))
    x = cloud(3, 77)
    cases["small"] = (0.3, 4, x, queries(x, 20))
    return cases


@pytest.mark.parametrize("name", sorted(_grid_cases()))
def test_ball_query_grid_matches_oracle(ops, name):
    _, _, _, orc = ops
    radius, ns, x1, x2 = _grid_cases()[name]
    idx, cnt = _ball_grid(radius, ns, to_cuda(x1), to_cuda(x2))
    eidx, ecnt = orc.query_ball_point(radius, ns, x1, x2)
    np.testing.assert_array_equal(cnt.cpu().numpy(), ecnt)
    np.testing.assert_array_equal(idx.cpu().numpy(), eidx)


def test_ball_query_grid_non_finite_inputs_match_default_kernel(ops):
    """A NaN distance is a hit in the reference (CUDA max(NaN,1e-20f) = 1e-20f): non-finite data or queries
    must take the brute-force path and agree with the default kernel."""
    _, tg, _, _ = ops
    rs = np.random.RandomState(3)
    x = rs.random_sample((2, 2048, 3)).astype(np.float32)
    q = x[:, :64].copy()
    q[0, 3, 1] = np.nan
    a, b = to_cuda(x), to_cuda(q)
    i0, c0 = tg.query_ball_point(0.2, 16, a, b)
    i1, c1 = _ball_grid(0.2, 16, a, b)
    assert bool((i0 == i1).all()) and bool((c0 == c1).all())
    x[1, 100, 2] = np.inf
    a = to_cuda(x)
    i0, c0 = tg.query_ball_point(0.2, 16, a, b)
    i1, c1 = _ball_grid(0.2, 16, a, b)
    assert bool((i0 == i1).all()) and bool((c0 == c1).all())


# ------------------------------------------------------------------ label vote with the fp32 gate
def test_knn_vote_filtered_matches_oracle_in_subprocess(cuda):
    """PN2_KNN_VOTE_FILTER is read once per process, so the gated kernel is exercised in a child process:
    labels and colours for knn 1..32 on uniform, lattice (ties) and large-offset clouds must equal the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import pn2_b200
from pn2_b200.tf_ops import tf_interpolate as ti
from oracle import oracle as orc
rs = np.random.RandomState(0)
cases = [(rs.random_sample((3000, 3)), rs.random_sample((2000, 3))),
         (rs.randint(0, 5, (900, 3)), rs.randint(0, 5, (700, 3))),
         (1000 + 1e-3 * rs.random_sample((2100, 3)), 1000 + 1e-3 * rs.random_sample((500, 3))),
         (rs.random_sample((2, 3)), rs.random_sample((50, 3)))]
for sp, dp in cases:
    sp, dp = sp.astype(np.float32), dp.astype(np.float32)
    sl = rs.randint(0, 9, len(sp)).astype(np.int32)
    for k in (1, 3, 5, 9, 17, 32):
        lab, col = ti.interpolate_label_with_color(torch.as_tensor(sp).cuda(), torch.as_tensor(sl).cuda(),
                                                   torch.as_tensor(dp).cuda(), k)
        el, ec = orc.interpolate_label_with_color(sp, sl, dp, k)
        assert (lab.cpu().numpy() == el).all() and (col.cpu().numpy() == ec).all(), (len(sp), k)
print("VOTE_FILTER_OK")
""" % root
    env = dict(os.environ, PN2_KNN_VOTE_FILTER="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert "VOTE_FILTER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
