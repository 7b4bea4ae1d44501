"""Diagnostic: decode what the tensor-core wgrad kernel computes (layout probe)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
ffi = pn2_b200._ffi
p = ffi.ptr

def wgrad(A, dY, mode):
    M, K = A.shape; N = dY.shape[1]
    At, dYt = torch.as_tensor(A).cuda(), torch.as_tensor(dY).cuda()
    dW = torch.zeros((K, N), dtype=torch.float32, device="cuda")
    ffi.call("pn2_linear_wgrad", M, K, N, p(At), K, None, None, 0, p(dYt), p(dW), None, mode)
    torch.cuda.synchronize()
    return dW.cpu().numpy()

M, K, N = 2048, 64, 64
rs = np.random.RandomState(0)
A = rs.normal(size=(M, K)).astype(np.float32); dY = rs.normal(size=(M, N)).astype(np.float32)
exp = A.astype(np.float64).T @ dY.astype(np.float64)
got = wgrad(A, dY, 1)
print("random: |got| max %.4g mean %.4g ; |exp| max %.4g ; corr(got,exp) %.4f ; corr(got,exp.T) %.4f" % (
    np.abs(got).max(), np.abs(got).mean(), np.abs(exp).max(),
    np.corrcoef(got.ravel(), exp.ravel())[0, 1] if got.std() > 0 else 0,
    np.corrcoef(got.ravel(), exp.T.ravel())[0, 1] if got.std() > 0 else 0))
print("got[:3,:6]", got[:3, :6]); print("exp[:3,:6]", exp[:3, :6])
# layout probe: A has ones in column k0 only for rows in [r0, r1); dY ones in column n0
for (k0, n0, r0, r1) in [(3, 5, 0, 2048), (3, 5, 0, 8), (3, 5, 8, 16), (40, 33, 0, 2048), (0, 0, 0, 1), (1, 0, 1, 2), (0, 1, 32, 33)]:
    A = np.zeros((M, K), np.float32); dY = np.zeros((M, N), np.float32)
    A[r0:r1, k0] = 1; dY[r0:r1, n0] = 1
    got = wgrad(A, dY, 1)
    nz = np.argwhere(np.abs(got) > 1e-6)
    print("probe k0=%d n0=%d rows[%d,%d): expected dW[%d,%d]=%d ; nonzeros: %s" % (
        k0, n0, r0, r1, k0, n0, r1 - r0, [(int(a), int(b), float(got[a, b])) for a, b in nz[:8]]))
