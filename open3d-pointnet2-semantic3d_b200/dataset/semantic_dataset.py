"""GPU-resident z-column box sampling: the reference's ``dataset/semantic_dataset.py`` sampling step.

  SemanticFileData(points, labels, colors, box_size_x, box_size_y)    semantic_dataset.py:58-88
      .sample_batch(batch_size, num_points_per_sample, ...)           semantic_dataset.py:150-216
                                                                      (+ provider.py:72-102 when augment)

The reference loads a scene with Open3D (absent here: the arrays are passed in), keeps it x-sorted on the
host and cuts one sample at a time with numpy; every training step then copies the batch to the device
(train.py:225-244).  Here the scene is uploaded ONCE (fp64 like Open3D's arrays) and a single kernel
launch (csrc/pn2_feed.cu) cuts the whole batch straight into the (B, N, 3+feat) float32 tensor that
``model.get_model`` consumes -- the per-step host feed disappears.

Randomness: centre indices and rotation angles come from the caller's ``rng`` (numpy's global stream by
default, the calls the reference makes: ``randint(0, P)`` per sample, ``uniform() * 2 pi`` per sample);
the random subset of an over-full box is drawn on the device from ``seed`` (see pn2_feed.cu), not from
numpy's shuffle.  There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from .._ffi import F32, F64, I32, call, ptr


class SemanticFileData:
    def __init__(self, points, labels, colors, box_size_x, box_size_y, device="cuda"):
        points = np.asarray(points, np.float64)
        if points.ndim != 2 or points.shape[1] != 3 or len(points) == 0:
            raise ValueError("points must be (num_points, 3)")
        n = len(points)
        labels = np.zeros(n, np.int32) if labels is None else np.asarray(labels).astype(np.int32)
        colors = None if colors is None else np.asarray(colors, np.float64)
        if labels.shape != (n,) or (colors is not None and (colors.ndim != 2 or len(colors) != n)):
            raise ValueError("labels must be (num_points,), colors (num_points, feat)")
        # semantic_dataset.py:84-88: sort by x to speed up the box extraction
        order = np.argsort(points[:, 0])
        self.box_size_x, self.box_size_y = float(box_size_x), float(box_size_y)
        self.sort_idx = order
        self.device = torch.device(device)
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dt)  # noqa: E731
        self.points = up(points[order], F64)
        self.labels = up(labels[order], I32)
        self.colors = None if colors is None else up(colors[order], F64)
        self.feat = 0 if colors is None else int(colors.shape[1])
        z = points[:, 2]
        self.scene_z_size = float(z.max() - z.min())  # semantic_dataset.py:131

    def __len__(self):
        return int(self.points.shape[0])

    def sample_batch(self, batch_size, num_points_per_sample, label_weights=None, augment=False,
                     rng=np.random, seed=0, center_idx=None, angles=None):
        """-> (data (B,N,3+feat) float32, labels (B,N) int32, weights (B,N) float32, index (B,N) int32
        scene indices of the x-sorted scene, count (B) points found per box); all CUDA tensors."""
        b, num, p = int(batch_size), int(num_points_per_sample), len(self)
        if b <= 0 or num <= 0:
            raise ValueError("batch_size and num_points_per_sample must be positive")
        if center_idx is None:  # semantic_dataset.py:154: points[np.random.randint(0, len(points))]
            center_idx = np.array([rng.randint(0, p) for _ in range(b)], np.int64)
        center_idx = np.asarray(center_idx, np.int64)
        if center_idx.shape != (b,) or center_idx.min() < 0 or center_idx.max() >= p:
            raise ValueError("center_idx must hold batch_size indices into the scene")
        if augment and angles is None:  # provider.py:83
            angles = np.array([rng.uniform() * 2 * np.pi for _ in range(b)], np.float64)
        dev = self.device
        d_center = torch.as_tensor(center_idx).to(dev)
        d_angles = None if angles is None else torch.as_tensor(np.asarray(angles, np.float64)).to(dev)
        d_lw, ncls = None, 0
        if label_weights is not None:
            d_lw = torch.as_tensor(np.asarray(label_weights, np.float32)).to(dev)
            ncls = int(d_lw.numel())
        data = torch.empty((b, num, 3 + self.feat), dtype=F32, device=dev)
        labels = torch.empty((b, num), dtype=I32, device=dev)
        weights = torch.empty((b, num), dtype=F32, device=dev)
        index = torch.empty((b, num), dtype=I32, device=dev)
        count = torch.empty((b,), dtype=I32, device=dev)
        call("pn2_box_sample", b, p, num, self.feat, ptr(self.points, F64), ptr(self.colors, F64, True),
             ptr(self.labels, I32), ptr(d_lw, F32, True), ncls, ptr(d_center, torch.int64),
             ptr(d_angles, F64, True), ctypes.c_double(self.box_size_x), ctypes.c_double(self.box_size_y),
             ctypes.c_double(self.scene_z_size), int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(data, F32),
             ptr(labels, I32), ptr(weights, F32), ptr(index, I32), ptr(count, I32))
        return data, labels, weights, index, count
