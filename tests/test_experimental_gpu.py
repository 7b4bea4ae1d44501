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
