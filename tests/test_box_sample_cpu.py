"""CPU: the numpy restatement of the reference's box sampling (oracle/box_sample_ref.py) -- its structure
against a literal transcription of the reference's control flow, and its key function against the library's
own host-callable copy (pn2_box_sample_key), so the GPU parity test compares like with like."""
import numpy as np

from oracle import box_sample_ref as bs


def scene(seed, n):
    rs = np.random.RandomState(seed)
    pts = rs.random_sample((n, 3)) * [40.0, 30.0, 6.0]
    pts = pts[np.argsort(pts[:, 0])]
    return pts, rs.randint(0, 9, n), rs.random_sample((n, 3))


def test_key_function_matches_the_library():
    import pn2_b200
    lib = pn2_b200._ffi.lib()
    idx = np.array([0, 1, 2, 77, 123456, 2**31 - 2], np.int64)
    for seed in (0, 1, 0xDEADBEEF12345678):
        for s in (0, 3, 15):
            exp = np.array([lib.pn2_box_sample_key(seed, s, int(i)) for i in idx], np.uint32)
            np.testing.assert_array_equal(bs.box_key(seed, s, idx), exp)


def test_restatement_follows_the_reference_control_flow():
    """semantic_dataset.py:150-186 transcribed literally (its own boolean masks), with the subset mask built from
    the documented key rule in place of np.random.shuffle: same points, same order, same centring."""
    pts, labels, colors = scene(3, 20000)
    for center, num in ((100, 512), (19999, 64), (7000, 4096)):
        out, lab, w, chosen, cnt = bs.sample(pts, labels, colors, center, num, 10.0, 10.0, 42, 1)
        mask = bs.extract_z_box(pts, pts[center], 10.0, 10.0)
        p_in, l_in = pts[mask], labels[mask]
        assert cnt == mask.sum()
        if len(p_in) - num > 0:
            keys = bs.box_key(42, 1, np.nonzero(mask)[0])
            kth = np.sort(keys)[num - 1]
            sample_mask = keys <= kth          # no ties in this data
            assert sample_mask.sum() == num
        else:
            sample_mask = np.arange(len(p_in))
            while len(sample_mask) < num:
                sample_mask = np.concatenate((sample_mask, sample_mask), axis=0)
            sample_mask = sample_mask[:num]
        p_s, l_s = p_in[sample_mask], l_in[sample_mask]
        box_min = np.min(p_s, axis=0)
        centered = p_s - np.array([box_min[0] + 5.0, box_min[1] + 5.0, box_min[2]])
        np.testing.assert_array_equal(out[:, :3], centered.astype(np.float32))
        np.testing.assert_array_equal(lab, l_s)
        assert out[:, 2].min() == 0.0 and out[:, 0].min() == -5.0
    # a sparse box is tiled: period = number of points in the box
    sparse, sl, sc = scene(4, 300)
    out, lab, w, chosen, cnt = bs.sample(sparse, sl, sc, 10, 256, 10.0, 10.0, 0, 0)
    assert cnt < 256
    assert np.array_equal(chosen, np.tile(chosen[:cnt], 256 // cnt + 1)[:256])
    # rotation about z keeps z and the xy norm
    o2, _, _, _, _ = bs.sample(pts, labels, colors, 100, 512, 10.0, 10.0, 42, 1, angle=0.7)
    o1, _, _, _, _ = bs.sample(pts, labels, colors, 100, 512, 10.0, 10.0, 42, 1)
    np.testing.assert_array_equal(o1[:, 2:], o2[:, 2:])
    np.testing.assert_allclose(np.hypot(o1[:, 0], o1[:, 1]), np.hypot(o2[:, 0], o2[:, 1]), rtol=1e-6)
