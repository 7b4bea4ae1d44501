"""GPU: z-column box sampling kernel (csrc/pn2_feed.cu) against the numpy
restatement of dataset/semantic_dataset.py:90-186 + provider.py:72-102 with the same centre indices, angles
and subset seed."""
import numpy as np
import pytest

from _util import to_cuda

pytestmark = pytest.mark.gpu


def scene(seed, n, extent=(40.0, 30.0, 6.0), dup=False):
    rs = np.random.RandomState(seed)
    pts = rs.random_sample((n, 3)) * list(extent)
    if dup:  # many points share an x value: searchsorted edges and the >= / <= tests see exact ties
        pts[:, 0] = np.round(pts[:, 0] * 2) / 2
    labels = rs.randint(0, 9, n).astype(np.int32)
    colors = rs.random_sample((n, 3))
    return pts, labels, colors


def run_case(pts, labels, colors, b, num, box, seed, augment, rs, lw=None):
    import pn2_b200  # noqa: F401
    from pn2_b200.dataset.semantic_dataset import SemanticFileData
    from oracle import box_sample_ref as bs
    fd = SemanticFileData(pts, labels, colors, box, box)
    order = fd.sort_idx
    sp, sl = pts[order], labels[order]
    sc = None if colors is None else colors[order]
    centers = rs.randint(0, len(pts), b)
    angles = rs.uniform(size=b) * 2 * np.pi if augment else None
    data, lab, w, index, count = fd.sample_batch(b, num, label_weights=lw, augment=augment, seed=seed,
                                                 center_idx=centers, angles=angles)
    data, lab, w, index, count = (t.cpu().numpy() for t in (data, lab, w, index, count))
    for i in range(b):
        e_out, e_lab, e_w, e_idx, e_cnt = bs.sample(sp, sl, sc, int(centers[i]), num, box, box, seed, i,
                                                    angle=None if angles is None else angles[i],
                                                    label_weights=lw)
        assert count[i] == e_cnt
        np.testing.assert_array_equal(index[i], e_idx)
        np.testing.assert_array_equal(lab[i], e_lab)
        np.testing.assert_array_equal(w[i], e_w)
        if augment:  # fp64 3-term dot products: BLAS / nvcc may or may not fuse -> last fp32 bit
            np.testing.assert_allclose(data[i, :, :3], e_out[:, :3], rtol=3e-7, atol=1e-6)
            np.testing.assert_array_equal(data[i, :, 3:], e_out[:, 3:])
        else:
            np.testing.assert_array_equal(data[i], e_out)
    return count


@pytest.mark.parametrize("n,b,num,box,augment,dup", [(200000, 16, 8192, 10.0, False, False),
                                                      (200000, 8, 8192, 10.0, True, False),
                                                      (50000, 4, 8192, 10.0, False, False),   # boxes tiled (short)
                                                      (30000, 6, 1000, 6.0, True, True),
                                                      (5000, 3, 8192, 100.0, False, False),   # whole scene, tiled
                                                      (3000, 2, 64, 2.0, False, True)])
def test_box_sample_matches_numpy_restatement(cuda, n, b, num, box, augment, dup):
    pts, labels, colors = scene(n, n, dup=dup)
    rs = np.random.RandomState(n + b)
    lw = rs.uniform(0.5, 2.0, 9).astype(np.float32)
    count = run_case(pts, labels, colors, b, num, box, 1234 + n, augment, rs, lw)
    assert (count > 0).all()


def test_box_sample_without_colors_and_labels_feeds_the_model(cuda):
    """use_color = 0 scenes (semantic_no_color.json): 3-channel output; the batch goes straight into
    model.get_model without touching the host."""
    import pn2_b200  # noqa: F401
    from pn2_b200 import model
    from pn2_b200.dataset.semantic_dataset import SemanticFileData
    from pn2_b200.util import tf_util
    pts, _, _ = scene(5, 60000)
    rs = np.random.RandomState(1)
    run_case(pts, np.zeros(len(pts), np.int32), None, 2, 1024, 10.0, 7, False, rs)
    fd = SemanticFileData(pts, None, None, 10.0, 10.0)
    data, lab, w, _, _ = fd.sample_batch(2, 1024, rng=np.random.RandomState(0), seed=3)
    assert tuple(data.shape) == (2, 1024, 3) and bool((w == 1).all())
    hp = {"use_color": 0, "l1_npoint": 256, "l1_radius": 0.5, "l1_nsample": 32, "l2_npoint": 64, "l2_radius": 1.0,
          "l2_nsample": 32, "l3_npoint": 16, "l3_radius": 2.0, "l3_nsample": 32, "l4_npoint": 8, "l4_radius": 4.0,
          "l4_nsample": 32}
    tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
    pred, _ = model.get_model(data, True, 9, hp, bn_decay=0.5)
    assert tuple(pred.shape) == (2, 1024, 9) and bool(pred.isfinite().all())


def test_box_sample_validation(cuda):
    import pn2_b200  # noqa: F401
    from pn2_b200.dataset.semantic_dataset import SemanticFileData
    pts, labels, colors = scene(1, 1000)
    with pytest.raises(ValueError, match="points must be"):
        SemanticFileData(pts[:, :2], labels, colors, 10, 10)
    fd = SemanticFileData(pts, labels, colors, 10.0, 10.0)
    with pytest.raises(ValueError, match="positive"):
        fd.sample_batch(0, 8192)
    with pytest.raises(ValueError, match="center_idx"):
        fd.sample_batch(2, 64, center_idx=np.array([0, 1000]))
