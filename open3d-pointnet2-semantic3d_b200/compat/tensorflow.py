"""The handful of ``tensorflow`` symbols the reference's model.py touches, eagerly, on torch tensors.

The north star asks that ``model.py`` "drops in unchanged".  /root/reference/model.py imports
``tensorflow`` (model.py:7) and uses exactly:

    tf.placeholder, tf.float32, tf.int32                 model.py:14-18
    tf.slice                                             model.py:28-29
    tf.losses.sparse_softmax_cross_entropy               model.py:156-158
    tf.summary.scalar, tf.add_to_collection              model.py:159-160

everything else goes through ``util.tf_util`` / ``util.pointnet_util`` (model.py:8-9), which this package
provides under the same names.  ``install()`` registers this module as ``tensorflow`` -- only when no real
TensorFlow is importable -- together with the reference's module names (``util.tf_util`` ...); after that
``import model`` of the UNMODIFIED reference file builds and runs the network on the sm_100a kernels.
There is no graph: placeholders are shape/dtype records, ops run eagerly, "collections" and "summaries"
are plain dictionaries.
"""
import collections
import sys
import types

import torch

float32 = torch.float32
int32 = torch.int32

Placeholder = collections.namedtuple("Placeholder", "dtype shape name")
_collections = collections.defaultdict(list)
_summaries = {}


def placeholder(dtype, shape=None, name=None):
    """tf.placeholder: there is nothing to feed -- a (dtype, shape) record (model.py:14-18)."""
    return Placeholder(dtype, tuple(shape) if shape is not None else None, name)


def slice(input_, begin, size, name=None):  # noqa: A001 - TensorFlow's name
    """tf.slice(t, begin, size): size -1 = to the end of that dimension (model.py:28-29)."""
    if len(begin) != input_.dim() or len(size) != input_.dim():
        raise ValueError("slice: begin/size must have one entry per dimension")
    idx = tuple(builtins_slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return input_[idx].contiguous()


builtins_slice = __builtins__["slice"] if isinstance(__builtins__, dict) else __builtins__.slice


def add_to_collection(name, value):
    _collections[name].append(value)


def get_collection(name):
    return list(_collections[name])


def _sparse_softmax_cross_entropy(labels, logits, weights=1.0, scope=None):
    """tf.losses.sparse_softmax_cross_entropy, reduction SUM_BY_NONZERO_WEIGHTS (model.py:156-158)."""
    from .. import model as _model
    if not isinstance(weights, torch.Tensor):
        weights = torch.full(tuple(labels.shape), float(weights), dtype=torch.float32, device=logits.device)
    return _model.get_loss(logits, labels, weights)


def _scalar(name, tensor):
    _summaries[name] = tensor
    return name


losses = types.SimpleNamespace(sparse_softmax_cross_entropy=_sparse_softmax_cross_entropy)
summary = types.SimpleNamespace(scalar=_scalar, values=_summaries)


def install():
    """Make ``import tensorflow`` resolve to this shim (never over a real TensorFlow) and register the
    reference's module names.  Returns True if the shim was installed."""
    import importlib.util
    from .. import install_reference_aliases
    install_reference_aliases(include_model=False)
    if "tensorflow" in sys.modules and sys.modules["tensorflow"] is not sys.modules[__name__]:
        return False
    try:
        if importlib.util.find_spec("tensorflow") is not None:
            return False
    except (ImportError, ValueError):
        pass
    sys.modules["tensorflow"] = sys.modules[__name__]
    return True
