"""CPU: the C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pn2_b200.h")
LIB = os.path.join(ROOT, "open3d-pointnet2-semantic3d_b200", "lib", "libpn2_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    if not os.path.exists(LIB):
        g.build()
    lib = ctypes.CDLL(LIB)
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert lib.pn2_abi_version() == 1


def test_python_binding_covers_header():
    import pn2_b200
    sig = set(pn2_b200._ffi.SIGNATURES)
    for n in declared_functions():
        assert n in sig, "no ctypes signature for " + n
    pn2_b200._ffi.lib()  # resolves all of them


def test_validation_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on CPU."""
    lib = ctypes.CDLL(LIB)
    lib.pn2_query_ball_point.argtypes = [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 5
    assert lib.pn2_query_ball_point(1, 8, 8, ctypes.c_float(-1.0), 4, None, None, None, None, None) == -1
    assert lib.pn2_query_ball_point(1, 8, 8, ctypes.c_float(0.5), 0, None, None, None, None, None) == -1
    assert lib.pn2_query_ball_point(1, 8, 8, ctypes.c_float(0.5), 4, None, None, None, None, None) == -4
    lib.pn2_fps.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4
    assert lib.pn2_fps(1, 8, 0, None, None, None, None) == -1  # "expects positive npoint"
    lib.pn2_strerror.restype = ctypes.c_char_p
    assert b"invalid" in lib.pn2_strerror(-1)


def test_ball_threshold_is_exact():
    """pn2_ball_threshold(r) = min{t : sqrtf(t) >= r}: so (sqrtf(d2) < r) == (d2 < T) for all d2."""
    import numpy as np
    lib = ctypes.CDLL(LIB)
    lib.pn2_ball_threshold.restype = ctypes.c_float
    lib.pn2_ball_threshold.argtypes = [ctypes.c_float]
    rs = np.random.RandomState(0)
    radii = np.concatenate([rs.uniform(1e-3, 8, 300), [0.1, 0.2, 0.4, 0.5, 1.0, 2.0, 4.0]]).astype(np.float32)
    for r in radii:
        t = np.float32(lib.pn2_ball_threshold(ctypes.c_float(float(r))))
        assert np.sqrt(t, dtype=np.float32) >= r
        below = np.nextafter(t, np.float32(0))
        assert np.sqrt(below, dtype=np.float32) < r
    assert lib.pn2_ball_threshold(ctypes.c_float(1e-21)) == 0.0
