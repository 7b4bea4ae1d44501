"""CPU: the oracle's prefix sum / prob_sample against an independent simulation of the reference's
shared-memory scan (tf_ops/tf_sampling.cu:7-92) and against the properties the op must have.

The reference has no golden vector for ProbSample (its only use is the smoke test
tf_ops/test_tf_ops.py:96-128, random numbers from tf.random_uniform), so the rounding ORDER of the
fp32 prefix sum is pinned by source: `scan_blockwise_sim` below walks the same buffers with the same
index arithmetic as the device code (quads -> padded tree up-sweep -> down-sweep -> carry), one fp32
addition at a time, while oracle/pn2_oracle.c states the same sum as a recurrence.  Both must agree
to the bit; on the GPU box the reference kernel itself is the third witness (test_ops_gpu.py).
"""
import numpy as np
import pytest

from oracle import oracle as orc

f32 = np.float32


def scan_blockwise_sim(row):
    """fp32 prefix sum of one row, simulating tf_sampling.cu:7-92 buffer by buffer."""
    n = len(row)
    out = np.zeros(n, f32)
    block, pad = 2048, 5
    run, run2 = f32(0), f32(0)
    for j in range(0, n, block * 4):
        n24_i = min(n - j, block * 4)
        n24 = (n24_i + 3) & ~3
        n2 = n24 >> 2
        buf4 = np.zeros(block * 4, f32)
        buf = np.zeros(block + (block >> pad), f32)
        at = lambda q: q + (q >> pad)  # noqa: E731  (bank-conflict padding of the tree buffer)
        for k in range(0, n24_i, 4):
            if k + 3 < n24_i:
                v1, v2, v3, v4 = (f32(row[j + k + i]) for i in range(4))
                v2 = f32(v2 + v1)
                v4 = f32(v4 + v3)
                v3 = f32(v3 + v2)
                v4 = f32(v4 + v2)
                buf4[k:k + 4] = (v1, v2, v3, v4)
                buf[at(k >> 2)] = v4
            else:
                v = f32(0)
                for k2 in range(k, n24_i):
                    v = f32(v + row[j + k2])
                    buf4[k2] = v
                buf4[n24_i:n24] = v
                buf[at(k >> 2)] = v
        u = 0
        while (2 << u) <= n2:
            for k in range(n2 >> (u + 1)):
                i1 = (((k << 1) + 2) << u) - 1
                i2 = (((k << 1) + 1) << u) - 1
                buf[at(i1)] = f32(buf[at(i1)] + buf[at(i2)])
            u += 1
        u -= 1
        while u >= 0:
            for k in range(max(0, (n2 - (1 << u)) >> (u + 1))):
                i1 = (((k << 1) + 3) << u) - 1
                i2 = (((k << 1) + 2) << u) - 1
                buf[at(i1)] = f32(buf[at(i1)] + buf[at(i2)])
            u -= 1
        for k in range(4, n24, 4):
            buf4[k:k + 4] = buf4[k:k + 4] + buf[at((k >> 2) - 1)]
        out[j:j + n24_i] = buf4[:n24_i] + run
        t = f32(buf[at(n2 - 1)] + run2)
        r2 = f32(run + t)
        run2 = f32(t - f32(r2 - run))
        run = r2
    return out


def search_sim(cdf, r):
    """tf_sampling.cu:94-110 for one row."""
    n = len(cdf)
    base = 1
    while base < n:
        base <<= 1
    res = np.empty(len(r), np.int32)
    for j, x in enumerate(r):
        q = f32(f32(x) * cdf[n - 1])
        pos = n - 1
        k = base
        while k >= 1:
            if pos >= k and cdf[pos - k] >= q:
                pos -= k
            k >>= 1
        res[j] = pos
    return res


SIZES = [1, 2, 3, 4, 5, 7, 8, 9, 63, 64, 65, 100, 1023, 1024, 1025, 4099, 8191, 8192, 8193, 8200,
         12345, 16384, 16389, 20000]


@pytest.mark.parametrize("n", SIZES)
def test_cumsum_rounding_order_matches_reference_scan(n):
    rs = np.random.RandomState(n)
    x = (rs.random_sample((2, n)) * rs.choice([1e-3, 1.0, 37.0], size=(2, n))).astype(f32)
    got = orc.cumsum(x)
    for i in range(2):
        exp = scan_blockwise_sim(x[i])
        np.testing.assert_array_equal(got[i].view(np.uint32), exp.view(np.uint32))


def test_cumsum_close_to_fp64():
    """The tree order keeps the error near one ulp of the total.  (The fp32 result is NOT guaranteed
    monotone -- neighbouring elements take different summation paths -- which is why the search
    below has to be the reference's exact descending-step search, not a generic lower bound.)"""
    rs = np.random.RandomState(3)
    x = rs.random_sample((3, 50000)).astype(f32)
    got = orc.cumsum(x)
    ref = np.cumsum(x.astype(np.float64), axis=1)
    assert np.abs(got - ref).max() <= 2e-7 * ref.max() * 4


@pytest.mark.parametrize("n,m", [(5, 8192), (1, 16), (33, 100), (8193, 500), (20000, 300)])
def test_prob_sample_matches_simulation(n, m):
    rs = np.random.RandomState(100 + n)
    p = rs.random_sample((2, n)).astype(f32)
    p[:, rs.randint(0, n, max(1, n // 7))] = 0  # zero-probability categories (flat CDF steps)
    r = rs.random_sample((2, m)).astype(f32)
    r[:, 0] = 0.0
    if m > 1:
        r[:, 1] = np.nextafter(f32(1), f32(0))
    got = orc.prob_sample(p, r)
    assert got.dtype == np.int32 and got.shape == (2, m)
    for i in range(2):
        cdf = scan_blockwise_sim(p[i])
        np.testing.assert_array_equal(got[i], search_sim(cdf, r[i]))


def test_prob_sample_is_inverse_cdf():
    """Defining property: result = first category whose cumulative weight reaches r * total, so a
    category with zero weight is never drawn (except category 0 for r == 0) and the draw
    frequencies follow the weights."""
    rs = np.random.RandomState(5)
    p = rs.random_sample((1, 40)).astype(f32)
    p[0, [3, 17, 18]] = 0
    r = rs.random_sample((1, 200000)).astype(f32)
    idx = orc.prob_sample(p, r)[0]
    assert idx.min() >= 0 and idx.max() < 40
    cdf = orc.cumsum(p)[0]
    q = r[0] * cdf[-1]
    assert (cdf[idx] >= q).all()
    prev = np.where(idx > 0, cdf[np.maximum(idx - 1, 0)], -1.0)
    assert (prev < q).all()
    counts = np.bincount(idx, minlength=40)
    assert counts[[3, 17, 18]].sum() == 0
    np.testing.assert_allclose(counts / counts.sum(), p[0] / p[0].sum(), atol=5e-3)


def test_prob_sample_like_reference_test():
    """tf_ops/test_tf_ops.py:96-128: triangle areas as weights, 8192 draws, gather, FPS 1024."""
    np.random.seed(100)
    tri = np.random.rand(1, 5, 3, 3).astype("float32")
    a, b, c = tri[:, :, 0], tri[:, :, 1], tri[:, :, 2]
    areas = np.sqrt((np.cross(b - a, c - a) ** 2).sum(2) + 1e-9).astype(f32)
    r = np.random.rand(1, 8192).astype(f32)
    ids = orc.prob_sample(areas, r)
    assert ids.shape == (1, 8192) and ids.min() >= 0 and ids.max() <= 4
    np.testing.assert_allclose(np.bincount(ids[0], minlength=5) / 8192.0, areas[0] / areas.sum(),
                               atol=0.02)
    us, vs = np.random.rand(1, 8192).astype(f32), np.random.rand(1, 8192).astype(f32)
    upv, umv = 1 - np.abs(us + vs - 1), us - vs
    us, vs = (upv + umv) * 0.5, (upv - umv) * 0.5
    ta, tb, tc = (orc.gather_point(t, ids) for t in (a, b, c))
    pts = (ta + (tb - ta) * us[..., None] + (tc - ta) * vs[..., None]).astype(f32)
    red = orc.gather_point(pts, orc.farthest_point_sample(1024, pts))
    assert red.shape == (1, 1024, 3) and np.isfinite(red).all()
