import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import pn2_b200
ffi = pn2_b200._ffi; p = ffi.ptr
M, K, N = 131072, 128, 128
A = torch.randn(M, K, device="cuda"); dY = torch.randn(M, N, device="cuda"); dW = torch.zeros(K, N, device="cuda")
def run(): ffi.call("pn2_linear_wgrad", M, K, N, p(A), K, None, None, 0, p(dY), p(dW), None, 1)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("PN2_DBG_WGRAD=%s : %.1f us" % (os.environ.get("PN2_DBG_WGRAD", "0"), e0.elapsed_time(e1) * 50))
