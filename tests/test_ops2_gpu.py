"""GPU parity of the two kernels promoted to default paths in round 2 after their tests ran green on a
B200 and an A/B timing was recorded (profiles/README_r02.md): the cluster FPS with the push + remote
mbarrier handshake (default of pn2_fps above 8192 points) and the hashed-grid ball query (default of
tf_grouping.query_ball_point at n >= 4096).  Both are also called through their explicit entry points."""
import numpy as np
import pytest

from _util import rng_cloud, to_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    import pn2_b200  # noqa: F401
    from pn2_b200.tf_ops import tf_grouping, tf_interpolate, tf_sampling
    from oracle import oracle as orc
    return tf_sampling, tf_grouping, tf_interpolate, orc


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
