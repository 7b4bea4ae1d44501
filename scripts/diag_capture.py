"""Diagnostic: which call invalidates the CUDA-graph capture of Trainer.forward_backward?
Wraps _ffi.call so that the capture status of the current stream is queried before/after every
entry point; prints the first call after which the capture is no longer active."""
import os, sys, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pn2_b200
from pn2_b200 import _ffi
from pn2_b200.train_step import Trainer

dev = torch.device("cuda", 0)
b, n = int(os.environ.get("B", 16)), 8192
pc, labels, smpw = bench.make_batch(b, n, 100)
d = [torch.as_tensor(x).to(dev) for x in (pc, labels, smpw)]
tr = Trainer(bench.HP, 9, device=dev, seed=0)
tr.step(*d); tr.step(*d)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
orig = _ffi.call
log = []
def status():
    try:
        r = rt.cudaStreamIsCapturing(torch.cuda.current_stream().cuda_stream)
        return r
    except Exception as e:
        return "exc:%r" % (e,)
def wrapped(name, *a):
    s0 = status()
    try:
        rc = orig(name, *a)
    except Exception as e:
        log.append((name, s0, "raise %r" % (e,)))
        raise
    s1 = status()
    log.append((name, s0, s1))
    return rc
ok = tr.capture(*d)
print("plain capture ->", ok, "|", getattr(tr, "_capture_error", None))
if not ok:
    # patch every module-level reference to call
    import pn2_b200.util.tf_util as tu, pn2_b200.util.pointnet_util as pu, pn2_b200.model as mo
    import pn2_b200.tf_ops.tf_sampling as ts, pn2_b200.tf_ops.tf_grouping as tg, pn2_b200.tf_ops.tf_interpolate as ti
    import pn2_b200.train_step as tst
    for m in (tu, pu, mo, ts, tg, ti, tst, _ffi):
        if hasattr(m, "call"):
            m.call = wrapped
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            tr.forward_backward(*d)
        print("second capture OK?!")
    except Exception:
        traceback.print_exc()
    prev = None
    for i, (name, s0, s1) in enumerate(log):
        if str(s0) != str(prev) or str(s1) != str(s0):
            print(i, name, s0, "->", s1)
        prev = s1
    print("calls logged", len(log))
