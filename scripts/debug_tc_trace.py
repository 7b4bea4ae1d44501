"""Diagnostic: cycle accounting of the tcgen05 forward kernel (CTA 0) on a few shapes."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
ffi = pn2_b200._ffi
p = ffi.ptr
lib = ffi.lib()
names = ["mma wait full", "mma issue", "mma wait acc_empty", "prod load+xform", "prod wait empty", "prod store",
         "epi wait acc_full", "epi process", "loader wait empty", "mma total", "chunks", "tiles",
         "epi wait store", "epi tmem->smem", "epi fence", "epi store issue"]
for (M, K, N, pro) in [(131072, 128, 128, 1), (131072, 128, 128, 0), (524288, 32, 32, 1), (8192, 256, 256, 1)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda") * 0.1
    Y = torch.empty(M, N, device="cuda"); sc = torch.ones(K, device="cuda"); sh = torch.zeros(K, device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    nb = int(lib.pn2_linear_workspace_bytes(K, N)); ws = torch.empty(max(nb // 4, 4), device="cuda")
    def run():
        ffi.call("pn2_linear_fwd", M, K, N, p(A), K, p(sc) if pro else None, p(sh) if pro else None, pro, p(W), None, p(Y),
                 p(stats) if pro else None, p(ws), ws.numel() * 4, 1)
    for _ in range(3): run()
    buf = (ctypes.c_longlong * 16)()
    lib.pn2_debug_tc_trace(buf)
    run()
    lib.pn2_debug_tc_trace(buf)
    v = list(buf)
    print("== M,K,N=%s prologue+stats=%d" % ((M, K, N), pro))
    ch = max(v[10], 1)
    for i, nme in enumerate(names):
        print("   %-20s %10d cycles  (%8.1f per chunk)" % (nme, v[i], v[i] / ch))

print("######## wgrad")
for (M, K, N) in [(131072, 128, 128), (131072, 64, 64), (524288, 32, 32), (4096, 256, 256)]:
    A = torch.randn(M, K, device="cuda"); dY = torch.randn(M, N, device="cuda"); dW = torch.zeros(K, N, device="cuda")
    sc = torch.ones(K, device="cuda"); sh = torch.zeros(K, device="cuda")
    def run():
        ffi.call("pn2_linear_wgrad", M, K, N, p(A), K, p(sc), p(sh), 1, p(dY), p(dW), None, 1)
    for _ in range(3): run()
    buf = (ctypes.c_longlong * 16)()
    lib.pn2_debug_tc_trace(buf); run(); lib.pn2_debug_tc_trace(buf)
    v = list(buf); ch = max(v[10], 1)
    print("== wgrad M,K,N=%s" % ((M, K, N),))
    names[6] = 'prod fence+arrive'
    for i in (0, 1, 2, 9, 10, 11):
        print("   %-20s %10d cycles  (%8.1f per stage)" % (names[i], v[i], v[i] / ch))
