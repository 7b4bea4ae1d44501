"""CPU: the product package never touches oracle/ and has no CPU fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "open3d-pointnet2-semantic3d_b200")


def product_files():
    for d, _, fs in os.walk(PKG):
        if os.sep + "lib" in d:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(d, f)
    yield os.path.join(ROOT, "pn2_b200.py")
    yield os.path.join(ROOT, "include", "pn2_b200.h")


def test_product_does_not_import_oracle():
    pat = re.compile(r"^\s*(from|import)\s+\.*oracle|liborc|_ref/", re.M)
    for p in product_files():
        src = open(p).read()
        assert not pat.search(src), p


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import pn2_b200
    ffi = pn2_b200._ffi
    monkeypatch.setattr(ffi, "_lib", None)
    monkeypatch.setattr(ffi, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        ffi.lib()
    except ffi.Pn2Error as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("missing library must raise")


def test_cpu_tensors_are_rejected():
    import torch
    import pn2_b200
    from pn2_b200.tf_ops import tf_sampling
    x = torch.rand(1, 16, 3)
    try:
        tf_sampling.farthest_point_sample(4, x)
    except pn2_b200._ffi.Pn2Error as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("CPU tensors must be rejected")
