"""Shared helpers for the parity tests."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rng_cloud(seed, b, n, scale=(1.0, 1.0, 1.0), shift=(0.0, 0.0, 0.0)):
    """np.random.seed(seed) legacy stream, uniform fp32 like the reference tests
    (tf_ops/test_tf_ops.py:12-15)."""
    rs = np.random.RandomState(seed)
    x = rs.random_sample((b, n, 3)).astype(np.float32)
    return (x * np.asarray(scale, np.float32) + np.asarray(shift, np.float32)).astype(np.float32)


def golden(name):
    """Committed fixture tests/golden/<name>.npz (written by tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def golden_inputs():
    """The seeded inputs make_golden.py froze the oracle outputs on (same draws, same order)."""
    np.random.seed(100)
    target = np.random.random((64, 8192, 3)).astype("float32")
    reference = np.random.random((64, 1024, 3)).astype("float32")
    inp = {"nn_q": np.ascontiguousarray(target[:1, :256]), "nn_known": np.ascontiguousarray(reference[:1])}
    del target, reference
    rs = np.random.RandomState(100)
    inp["xyz"] = rs.random_sample((2, 1024, 3)).astype(np.float32)
    np.random.seed(100)
    tri = np.random.rand(1, 5, 3, 3).astype("float32")
    ta, tb, tc = tri[:, :, 0], tri[:, :, 1], tri[:, :, 2]
    inp["areas"] = np.sqrt((np.cross(tb - ta, tc - ta) ** 2).sum(2) + 1e-9).astype(np.float32)
    inp["r"] = np.random.rand(1, 8192).astype(np.float32)
    rs = np.random.RandomState(100)
    inp["w"] = rs.random_sample((1, 9000)).astype(np.float32)
    inp["sp"] = rs.random_sample((700, 3)).astype(np.float32)
    inp["sl"] = rs.randint(0, 9, 700).astype(np.int32)
    inp["dp"] = rs.random_sample((400, 3)).astype(np.float32)
    return inp


def to_cuda(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


class RefKernels:
    """The reference's OWN CUDA kernels (oracle/_ref/libref_tfops.so, built from
    /root/reference/tf_ops/*.cu by oracle/Makefile).  Device pointers in, nothing copied."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "_ref", "libref_tfops.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def prob_sample(self, inp, inpr):
        """-> (indices, cdf): tf_sampling.cu:212-216 (cumsumKernel + binarysearchKernel)."""
        import torch
        b, n = inp.shape
        m = inpr.shape[1]
        temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
        out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_prob_sample(b, n, m, self._p(inp), self._p(inpr), self._p(temp),
                                      self._p(out), 1)
        assert rc == 0, rc
        return out, temp

    def fps(self, inp, m):
        import torch
        b, n, _ = inp.shape
        temp = torch.empty((32, n), dtype=torch.float32, device=inp.device)
        out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_fps(b, n, m, self._p(inp), self._p(temp), self._p(out), 1)
        assert rc == 0, rc
        return out

    def query_ball_point(self, radius, nsample, xyz1, xyz2):
        import torch
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz1.device)
        cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_query_ball_point(b, n, m, ctypes.c_float(radius), nsample, self._p(xyz1),
                                           self._p(xyz2), self._p(idx), self._p(cnt), 1)
        assert rc == 0, rc
        return idx, cnt

    def group_point(self, points, idx):
        import torch
        b, n, c = points.shape
        _, m, ns = idx.shape
        out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_group_point(b, n, c, m, ns, self._p(points), self._p(idx), self._p(out), 1)
        assert rc == 0, rc
        return out

    def gather_point(self, inp, idx):
        import torch
        b, n, _ = inp.shape
        m = idx.shape[1]
        out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_gather_point(b, n, m, self._p(inp), self._p(idx), self._p(out), 1)
        assert rc == 0, rc
        return out

    def selection_sort(self, k, dist):
        """-> (outi, out): the reference's own selection_sort_gpu (tf_grouping.cu:95-136)."""
        import torch
        b, m, n = dist.shape
        outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
        out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_selection_sort(b, n, m, int(k), self._p(dist), self._p(outi), self._p(out), 1)
        assert rc == 0, rc
        return outi, out

    def gather_point_grad(self, inp, idx, out_g):
        import torch
        b, n, _ = inp.shape
        m = idx.shape[1]
        g = torch.empty((b, n, 3), dtype=torch.float32, device=inp.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_gather_point_grad(b, n, m, self._p(out_g), self._p(idx), self._p(g), 1)
        assert rc == 0, rc
        return g

    def group_point_grad(self, points, idx, grad_out):
        import torch
        b, n, c = points.shape
        _, m, ns = idx.shape
        g = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
        torch.cuda.synchronize()
        rc = self.lib.ref_group_point_grad(b, n, c, m, ns, self._p(grad_out), self._p(idx), self._p(g), 1)
        assert rc == 0, rc
        return g
