"""Inference wrapper: the reference's ``Predictor`` (predict.py:15-105) on the sm_100a engine.

  Predictor(checkpoint, num_classes, hyper_params)
      .predict(batch_data)                  eval-mode forward + arg-max           predict.py:65-91
      .interpolate_labels(sparse_points, sparse_labels, dense_points, knn=3)      predict.py:93-105

``checkpoint`` is the ``variable name -> array`` dictionary of a TensorFlow checkpoint of this
network (the reference's scope names, see model.py), or the path of an ``.npz`` holding it; there is
no TF Saver here.  Everything runs on the GPU: the network through model.get_model, the dense-cloud
label transfer through tf_ops.tf_interpolate.interpolate_label_with_color (the reference does that
step on the host through an Open3D KD-tree).
"""
import numpy as np
import torch

from . import model
from .tf_ops.tf_interpolate import interpolate_label_with_color
from .util import tf_util


def _as_cuda(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dtype).contiguous()


class Predictor:
    def __init__(self, checkpoint, num_classes, hyper_params, device="cuda"):
        self.num_classes, self.hyper_params = int(num_classes), hyper_params
        self.device = torch.device(device)
        self.store = tf_util.VariableStore(device=device, seed=0)
        if isinstance(checkpoint, str):
            with np.load(checkpoint) as z:
                checkpoint = {k: z[k] for k in z.files}
        if not checkpoint:
            raise ValueError("Predictor needs the checkpoint's variables (name -> array)")
        self.skipped = self.store.load_state_dict(checkpoint)  # optimizer slots etc. of a full TF checkpoint
        # a variable the checkpoint lacks must fail loudly, not be xavier-initialised during predict()
        self.store.strict = True

    def predict(self, batch_data):
        """batch_data (batch_size, num_point, 3 or 6) -> labels (batch_size, num_point), the arg-max
        of the eval-mode logits (moving-average BatchNorm, no dropout)."""
        x = _as_cuda(batch_data, torch.float32, self.device)
        if x.dim() != 3 or x.shape[2] != 3 + 3 * int(self.hyper_params["use_color"]):
            raise ValueError("batch_data must be (batch_size, num_point, %d)"
                             % (3 + 3 * int(self.hyper_params["use_color"])))
        tf_util.set_default_store(self.store)
        with torch.no_grad():
            pred, _ = model.get_model(x, False, self.num_classes, self.hyper_params)
            labels = torch.argmax(pred, dim=2)
        return labels.cpu().numpy()

    def interpolate_labels(self, sparse_points, sparse_labels, dense_points, knn=3):
        """Labels and colours of the full-resolution cloud from the labelled sparse cloud
        (k nearest neighbours vote, predict.py:93-105 feeds knn=3)."""
        sp = _as_cuda(sparse_points, torch.float32, self.device)
        sl = _as_cuda(sparse_labels, torch.int32, self.device)
        dp = _as_cuda(dense_points, torch.float32, self.device)
        labels, colors = interpolate_label_with_color(sp, sl, dp, int(knn))
        return labels.cpu().numpy(), colors.cpu().numpy()
