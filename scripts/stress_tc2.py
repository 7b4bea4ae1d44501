"""Stress (diagnostic): Y = A * I with A[m,k] = (m % 4096) + k/256 (exact in 3xTF32), so every wrong
output element names the element it was copied from."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200 import _ffi as ffi
p = ffi.ptr
M, K, N = 40000, 128, 128
iters = int(os.environ.get("STRESS_ITERS", "200"))
m_idx = torch.arange(M, device="cuda", dtype=torch.float32) % 4096
A = (m_idx[:, None] + torch.arange(K, device="cuda", dtype=torch.float32)[None, :] / 256).contiguous()
W = torch.eye(K, N, device="cuda").contiguous()
ws = torch.empty(int(ffi.lib().pn2_linear_workspace_bytes(K, N)) // 4 + 4, device="cuda")
junk = torch.empty(1 << 26, device="cuda")
Ys = [torch.empty(M, N, device="cuda") for _ in range(8)]
nbad = 0
for i in range(iters):
    Y = Ys[i % 8]
    Y.fill_(-1.0)
    if i % 3 == 0: junk.fill_(float(i))
    ffi.call("pn2_linear_fwd", M, K, N, p(A), K, None, None, 0, p(W), None, p(Y), None, p(ws), ws.numel() * 4, 1)
    if i % 8 == 7 or i == iters - 1:
        torch.cuda.synchronize()
        for j, Yc in enumerate(Ys):
            bad = (Yc != A).nonzero()
            if bad.shape[0]:
                nbad += 1
                rows = bad[:, 0].unique()
                print("iter ~%d: %d bad elements, rows %s" % (i - 7 + j, bad.shape[0], rows[:12].tolist()))
                for (r, c) in bad[:: max(1, bad.shape[0] // 6)][:6].tolist():
                    v = Yc[r, c].item()
                    print("    Y[%d (tile %d row %d), col %d] = %.6f  -> source row%%4096 = %d (expected %d), col %d" % (
                        r, r // 128, r % 128, c, v, int(v), r % 4096, int(round((v - int(v)) * 256))))
print("BAD ITERATIONS", nbad, "of", iters)
