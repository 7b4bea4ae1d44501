"""numpy restatement of the reference's z-column box sampling.  TEST INFRASTRUCTURE ONLY.

Follows dataset/semantic_dataset.py (reference file:line):
  _extract_z_box              :123-148   (searchsorted over the x-sorted scene, >= / <= on all three axes)
  _get_fix_sized_sample_mask  : 90-107   (boolean mask keeps scene order; tiling by repeated doubling)
  _center_box                 :109-121
  sample                      :150-186
and util/provider.py rotate_feature_point_cloud :72-102 (fp64 rotation, result stored as float32).

The one deliberate difference: the random subset of an over-full box.  The reference shuffles a boolean
mask with numpy's global Mersenne Twister (:97-99); the device cannot replay that stream, so the product
keeps the ``num`` smallest of the keys box_key(seed, sample, scene index) (ties: lower scene index) --
restated here in numpy uint64 arithmetic.  PARITY UNPINNED beyond this restatement (the reference has no
test of the sampler and its module imports open3d, which is absent).
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def box_key(seed, sample, idx):
    """csrc/pn2_feed.cu box_key: splitmix64 finaliser, upper 32 bits."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.uint64(0xD1B54A32D192ED03) * np.uint64(sample + 1)
             + np.uint64(0x9E3779B97F4A7C15) * (np.asarray(idx).astype(np.uint64) + np.uint64(1)))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def extract_z_box(points, center_point, box_size_x, box_size_y):
    """semantic_dataset.py:123-148, verbatim in effect: boolean mask over the x-sorted scene."""
    scene_z_size = np.max(points, axis=0)[2] - np.min(points, axis=0)[2]
    box_min = center_point - [box_size_x / 2, box_size_y / 2, scene_z_size]
    box_max = center_point + [box_size_x / 2, box_size_y / 2, scene_z_size]
    i_min = np.searchsorted(points[:, 0], box_min[0])
    i_max = np.searchsorted(points[:, 0], box_max[0])
    mask = np.sum((points[i_min:i_max, :] >= box_min) * (points[i_min:i_max, :] <= box_max), axis=1) == 3
    return np.hstack((np.zeros(i_min, dtype=bool), mask, np.zeros(len(points) - i_max, dtype=bool)))


def sample(points, labels, colors, center_idx, num_points, box_size_x, box_size_y, seed, sample_no,
           angle=None, label_weights=None):
    """One training sample -> (data (num,3+feat) float32, labels, weights, scene indices, box count)."""
    points = np.asarray(points, np.float64)
    mask = extract_z_box(points, points[center_idx], box_size_x, box_size_y)
    in_box = np.nonzero(mask)[0]
    assert len(in_box) != 0
    if len(in_box) - num_points > 0:  # :95-99, subset chosen by key instead of np.random.shuffle
        keys = box_key(seed, sample_no, in_box)
        keep = np.lexsort((in_box, keys))[:num_points]
        chosen = np.sort(in_box[keep])       # a boolean mask keeps scene order
    else:                                    # :100-106
        sample_mask = np.arange(len(in_box))
        while len(sample_mask) < num_points:
            sample_mask = np.concatenate((sample_mask, sample_mask), axis=0)
        chosen = in_box[sample_mask[:num_points]]
    pts = points[chosen]
    box_min = np.min(pts, axis=0)            # :109-121
    shift = np.array([box_min[0] + box_size_x / 2, box_min[1] + box_size_y / 2, box_min[2]])
    centered = pts - shift
    feat = 0 if colors is None else colors.shape[1]
    out = np.zeros((num_points, 3 + feat), np.float32)
    if feat:
        out[:, 3:] = np.asarray(colors, np.float64)[chosen]
    if angle is None:
        out[:, :3] = centered
    else:                                    # provider.py:83-101, rotation_axis "z"
        cosval, sinval = np.cos(angle), np.sin(angle)
        rot = np.array([[cosval, sinval, 0], [-sinval, cosval, 0], [0, 0, 1]])
        out[:, :3] = np.dot(centered.reshape((-1, 3)), rot)
    lab = np.asarray(labels)[chosen].astype(np.int32)
    w = np.ones(num_points, np.float32) if label_weights is None else np.asarray(label_weights, np.float32)[lab]
    return out, lab, w, chosen.astype(np.int32), len(in_box)
