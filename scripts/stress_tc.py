"""Stress (diagnostic): many back-to-back tensor-core GEMM calls with a deep launch queue and a
freshly flushed L2, each compared with the fp32 kernel's result afterwards."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200 import _ffi as ffi
p = ffi.ptr
torch.manual_seed(0)
shapes = [(4096, 131, 128), (16384, 131, 128), (65536, 32, 32), (16384, 67, 64), (1024, 259, 256), (40000, 128, 128),
          (524288, 32, 32), (524288, 6, 32), (131072, 64, 32), (33000, 32, 9), (131072, 128, 128), (8192, 768, 256)]
iters = int(os.environ.get("STRESS_ITERS", "30"))
junk = torch.empty(1 << 26, device="cuda")
total_bad = 0
for (M, K, N) in shapes:
    lda = (K + 3) // 4 * 4
    A = torch.randn(M, lda, device="cuda"); A[:, K:] = float("nan")
    W = torch.randn(K, N, device="cuda") * 0.1; b = torch.randn(N, device="cuda")
    sc = torch.rand(K, device="cuda") + 0.5; sh = torch.rand(K, device="cuda") - 0.5
    ws = torch.empty(int(ffi.lib().pn2_linear_workspace_bytes(K, N)) // 4 + 4, device="cuda")
    def run(mode, Y, stats):
        ffi.call("pn2_linear_fwd", M, K, N, p(A), lda, p(sc), p(sh), 1, p(W), p(b), p(Y), p(stats), p(ws), ws.numel() * 4, mode)
    Y0 = torch.empty(M, N, device="cuda"); s0 = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    run(0, Y0, s0)
    Ys = [torch.empty(M, N, device="cuda") for _ in range(iters)]
    Ss = [torch.zeros(2 * N, dtype=torch.float64, device="cuda") for _ in range(iters)]
    torch.cuda.synchronize()
    for i in range(iters):
        if i % 3 == 0: junk.fill_(float(i))          # flush L2, keep the queue deep
        if i % 3 == 1: A.add_(0.0)                    # rewrite A right before the GEMM reads it
        run(1, Ys[i], Ss[i])
    torch.cuda.synchronize()
    bad = 0
    for i in range(iters):
        d = (Ys[i] - Y0).abs().max().item()
        ds = ((Ss[i] - s0).abs() / (s0.abs() + 1.0)).max().item()
        if not (d < 1e-4 and ds < 1e-3):
            bad += 1
            rows = ((Ys[i] - Y0).abs() > 1e-4).nonzero()[:, 0].unique()
            print("   iter %d: max err %.3g stats err %.3g bad rows %s" % (i, d, ds, rows[:10].tolist()))
    print("%s: %d / %d bad" % ((M, K, N), bad, iters))
    total_bad += bad
print("TOTAL BAD", total_bad)
