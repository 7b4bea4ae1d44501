"""Micro-benchmark (CUDA events) of the linear entry points on the model's shapes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
ffi = pn2_b200._ffi
p = ffi.ptr

def timeit(fn, iters=10):
    """GPU time per call: `iters` calls captured into one CUDA graph (no per-call CPU overhead)."""
    for _ in range(2): fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters

shapes = [(524288, 32, 32), (524288, 32, 64), (131072, 67, 64), (131072, 64, 128), (131072, 128, 128),
          (131072, 131, 128), (32768, 131, 128), (32768, 128, 256), (8192, 256, 256), (8192, 256, 512),
          (16384, 320, 256), (4096, 384, 256), (1024, 256, 256), (128, 128, 128), (2048, 128, 128)]
print("%-22s %9s %9s %9s %9s %9s %9s   hbm-bound(us) tensor-bound(us)" % ("M,K,N", "fwd tc", "fwd simt", "dgr tc", "dgr simt", "wgr tc", "wgr simt"))
for M, K, N in shapes:
    lda = (K + 3) // 4 * 4  # rows padded to 16 bytes like the layers' concat buffers
    A = torch.randn(M, lda, device="cuda"); W = torch.randn(K, N, device="cuda") * 0.1
    Y = torch.empty(M, N, device="cuda"); dY = torch.randn(M, N, device="cuda"); dX = torch.empty(M, lda, device="cuda")
    dW = torch.zeros(K, N, device="cuda"); sc = torch.ones(K, device="cuda"); sh = torch.zeros(K, device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    nb = int(ffi.lib().pn2_linear_workspace_bytes(K, N)); ws = torch.empty(max(nb // 4, 4), device="cuda")
    res = []
    for mode in (1, 0):
        res.append(timeit(lambda: ffi.call("pn2_linear_fwd", M, K, N, p(A), lda, p(sc), p(sh), 1, p(W), None, p(Y), p(stats), p(ws), ws.numel() * 4, mode)))
    for mode in (1, 0):
        res.append(timeit(lambda: ffi.call("pn2_linear_dgrad", M, K, N, p(dY), p(W), p(dX), lda, p(ws), ws.numel() * 4, mode)))
    for mode in (1, 0):
        if mode == 1 and M < 512:
            res.append(float("nan")); continue
        res.append(timeit(lambda: ffi.call("pn2_linear_wgrad", M, K, N, p(A), lda, p(sc), p(sh), 1, p(dY), p(dW), None, mode)))
    hbm = M * (K + N) * 4 / 6.5e12 * 1e6
    tens = 3 * M * K * N / 5.5e14 * 1e6
    print("%-22s %9.1f %9.1f %9.1f %9.1f %9.1f %9.1f   %8.1f %8.1f" % ((M, K, N), res[0], res[2 - 1], res[2], res[3], res[4], res[5], hbm, tens))
