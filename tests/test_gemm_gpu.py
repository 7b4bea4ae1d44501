"""GPU parity of the shared-MLP GEMM entry points (pn2_linear_fwd / dgrad / wgrad), called
straight through the C ABI: the exact fp32 CUDA-core kernel (mode 0) and the tcgen05 3xTF32
tensor-core kernel (mode 1) against an fp64 matmul.  Tolerance 1e-5 absolute on O(1) outputs."""
import numpy as np
import pytest

from _util import to_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffi(cuda):
    import pn2_b200
    return pn2_b200._ffi


def _ws(ffi, k, n):
    import torch
    nb = int(ffi.lib().pn2_linear_workspace_bytes(k, n))
    return torch.empty(max(nb // 4, 4), dtype=torch.float32, device="cuda")


def linear_fwd(ffi, A, W, bias, scale, shift, relu, mode, lda=None, want_stats=True):
    import torch
    M = A.shape[0]
    K, N = W.shape
    Y = torch.empty((M, N), dtype=torch.float32, device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if want_stats else None
    ws = _ws(ffi, K, N)
    p = ffi.ptr
    ffi.call("pn2_linear_fwd", M, K, N, p(A), lda or K, p(scale, None, True), p(shift, None, True),
             1 if relu else 0, p(W), p(bias, None, True), p(Y), p(stats, None, True), p(ws),
             ws.numel() * 4, mode)
    return Y, stats


SHAPES = [  # M, K, N
    (128, 32, 32), (256, 64, 64), (1000, 67, 64), (4096, 131, 128), (640, 259, 256),
    (512, 256, 512), (384, 128, 48), (130, 16, 16), (8192, 128, 128), (2048, 384, 256),
    (1024, 512, 128), (1024, 768, 256), (2000, 640, 384), (20000, 1024, 128), (3000, 544, 192),
    (30000, 96, 256),
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_fwd(ffi, mode, M, K, N):
    import torch
    rs = np.random.RandomState(M + K + N)
    A = rs.normal(size=(M, K)).astype(np.float32)
    W = (rs.uniform(-1, 1, (K, N)) * np.sqrt(6.0 / (K + N))).astype(np.float32)
    b = rs.uniform(-0.5, 0.5, N).astype(np.float32)
    sc = rs.uniform(0.5, 1.5, K).astype(np.float32)
    sh = rs.uniform(-0.5, 0.5, K).astype(np.float32)
    At, Wt = to_cuda(A), to_cuda(W)
    # plain
    Y, st = linear_fwd(ffi, At, Wt, to_cuda(b), None, None, False, mode)
    exp = A.astype(np.float64) @ W.astype(np.float64) + b
    np.testing.assert_allclose(Y.cpu().numpy(), exp, atol=1e-5)
    # statistics are sums over M of values that each carry ~1e-6 relative error
    np.testing.assert_allclose(st.cpu().numpy()[:N], exp.sum(0), rtol=1e-5, atol=1e-5 * M)
    np.testing.assert_allclose(st.cpu().numpy()[N:], (exp ** 2).sum(0), rtol=1e-5, atol=1e-5 * M)
    # with the previous layer's BN affine + ReLU applied on the fly
    Y2, _ = linear_fwd(ffi, At, Wt, None, to_cuda(sc), to_cuda(sh), True, mode, want_stats=False)
    A2 = np.maximum(A.astype(np.float64) * sc + sh, 0.0)
    np.testing.assert_allclose(Y2.cpu().numpy(), A2 @ W.astype(np.float64), atol=1e-5)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("M,K,N", [(65536, 32, 32), (65536, 32, 64), (16384, 67, 64), (4096, 131, 128),
                                   (1024, 259, 256), (128, 256, 256), (2048, 320, 256),
                                   (40000, 131, 128), (40000, 128, 128), (30000, 6, 32)])
def test_linear_fwd_padded_rows_many_tiles(ffi, mode, M, K, N):
    """The layers' shapes: A in a row-padded buffer (lda = K rounded up to 4, NaN in the padding),
    more row tiles than SMs (persistent CTAs loop), BN prologue + bias + statistics together."""
    rs = np.random.RandomState(M + K + N)
    A = rs.normal(size=(M, K)).astype(np.float32)
    W = (rs.uniform(-1, 1, (K, N)) * np.sqrt(6.0 / (K + N))).astype(np.float32)
    b = rs.uniform(-0.5, 0.5, N).astype(np.float32)
    sc = rs.uniform(0.5, 1.5, K).astype(np.float32)
    sh = rs.uniform(-0.5, 0.5, K).astype(np.float32)
    lda = (K + 3) // 4 * 4
    Apad = np.full((M, lda), np.nan, np.float32)
    Apad[:, :K] = A
    Y, st = linear_fwd(ffi, to_cuda(Apad), to_cuda(W), to_cuda(b), to_cuda(sc), to_cuda(sh), True,
                       mode, lda=lda)
    A2 = np.maximum(A.astype(np.float64) * sc + sh, 0.0)
    exp = A2 @ W.astype(np.float64) + b
    np.testing.assert_allclose(Y.cpu().numpy(), exp, atol=1e-5)
    np.testing.assert_allclose(st.cpu().numpy()[:N], exp.sum(0), rtol=1e-5, atol=1e-5 * M)
    np.testing.assert_allclose(st.cpu().numpy()[N:], (exp ** 2).sum(0), rtol=1e-5, atol=1e-5 * M)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("M,K,N", [(256, 64, 64), (1000, 67, 64), (4096, 131, 128),
                                   (512, 259, 256), (384, 512, 256), (8192, 128, 128),
                                   (40000, 128, 128), (50000, 64, 32), (1024, 768, 256),
                                   (2048, 384, 1024), (30000, 256, 128)])
def test_linear_dgrad(ffi, mode, M, K, N):
    import torch
    rs = np.random.RandomState(7 + M + K + N)
    dY = rs.normal(size=(M, N)).astype(np.float32)
    W = (rs.uniform(-1, 1, (K, N)) * np.sqrt(6.0 / (K + N))).astype(np.float32)
    dX = torch.empty((M, K), dtype=torch.float32, device="cuda")
    ws = _ws(ffi, K, N)
    p = ffi.ptr
    dYt, Wt = to_cuda(dY), to_cuda(W)  # keep the device buffers alive across the call
    ffi.call("pn2_linear_dgrad", M, K, N, p(dYt), p(Wt), p(dX), K, p(ws), ws.numel() * 4, mode)
    exp = dY.astype(np.float64) @ W.astype(np.float64).T
    np.testing.assert_allclose(dX.cpu().numpy(), exp, atol=1e-5)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("M,K,N", [(1000, 6, 32), (4096, 67, 64), (2048, 131, 128), (8192, 32, 32),
                                   (300, 259, 256), (5000, 6, 32), (16384, 128, 128),
                                   (4100, 259, 256), (3072, 320, 48), (70000, 64, 64),
                                   (600, 768, 256), (33000, 64, 128), (20000, 32, 64),
                                   (9000, 96, 96), (1024, 256, 512), (40000, 6, 32)])
def test_linear_wgrad(ffi, mode, M, K, N):
    """A lives in a row-padded buffer (lda = K rounded up to 4, NaN in the padding) the way the
    layers allocate it, so that the TMA tensor-map path is taken for odd K as well."""
    if mode == 1 and M < 512:
        pytest.skip("tensor-core wgrad needs M >= 512")
    import torch
    rs = np.random.RandomState(11 + M + K + N)
    A = rs.normal(size=(M, K)).astype(np.float32)
    dY = (rs.normal(size=(M, N)) * 0.1).astype(np.float32)
    sc = rs.uniform(0.5, 1.5, K).astype(np.float32)
    sh = rs.uniform(-0.5, 0.5, K).astype(np.float32)
    dW = torch.zeros((K, N), dtype=torch.float32, device="cuda")
    db = torch.zeros(N, dtype=torch.float32, device="cuda")
    p = ffi.ptr
    lda = (K + 3) // 4 * 4
    Apad = np.full((M, lda), np.nan, np.float32)
    Apad[:, :K] = A
    At, sct, sht, dYt = to_cuda(Apad), to_cuda(sc), to_cuda(sh), to_cuda(dY)
    ffi.call("pn2_linear_wgrad", M, K, N, p(At), lda, p(sct), p(sht), 1, p(dYt), p(dW), p(db), mode)
    A2 = np.maximum(A.astype(np.float64) * sc + sh, 0.0)
    exp = A2.T @ dY.astype(np.float64)
    tol = 2e-5 * max(1.0, np.abs(exp).max())
    np.testing.assert_allclose(dW.cpu().numpy(), exp, atol=tol)
    np.testing.assert_allclose(db.cpu().numpy(), dY.astype(np.float64).sum(0), atol=tol)


def test_linear_wgrad_unpadded_rows_fall_back(ffi):
    """Rows that are not 16-byte aligned (lda = 67) cannot be described by a tensor map: auto mode
    must silently use the fp32 kernel, forced tensor-core mode must say so."""
    import torch
    M, K, N = 2048, 67, 64
    rs = np.random.RandomState(5)
    A = rs.normal(size=(M, K)).astype(np.float32)
    dY = (rs.normal(size=(M, N)) * 0.1).astype(np.float32)
    p = ffi.ptr
    At, dYt = to_cuda(A), to_cuda(dY)
    dW = torch.zeros((K, N), dtype=torch.float32, device="cuda")
    ffi.call("pn2_linear_wgrad", M, K, N, p(At), K, None, None, 0, p(dYt), p(dW), None, -1)
    exp = A.astype(np.float64).T @ dY.astype(np.float64)
    np.testing.assert_allclose(dW.cpu().numpy(), exp, atol=2e-5 * max(1.0, np.abs(exp).max()))
    with pytest.raises(Exception):
        ffi.call("pn2_linear_wgrad", M, K, N, p(At), K, None, None, 0, p(dYt), p(dW), None, 1)


def test_three_tf32_beats_plain_tf32_bound(ffi):
    """The compensated product must be far more accurate than a single TF32 pass could be:
    with K=256 and O(1) operands a 10-bit mantissa gives ~1e-3 absolute error."""
    rs = np.random.RandomState(0)
    A = rs.normal(size=(1024, 256)).astype(np.float32)
    W = rs.normal(size=(256, 128)).astype(np.float32) * 0.05
    Y, _ = linear_fwd(ffi, to_cuda(A), to_cuda(W), None, None, None, False, 1, want_stats=False)
    err = np.abs(Y.cpu().numpy() - A.astype(np.float64) @ W.astype(np.float64)).max()
    assert err < 1e-5, err


@pytest.mark.parametrize("M,N,ldz,relu", [(5000, 32, 32, 1), (4096, 64, 64, 1), (1000, 9, 9, 0),
                                          (3000, 128, 160, 1), (257, 512, 512, 1), (70000, 32, 32, 1)])
def test_bn_backward_kernels(ffi, M, N, ldz, relu):
    """pn2_bn_train_finalize + pn2_bn_bwd_reduce + pn2_bn_bwd_apply (scalar and float4 paths) against
    the closed-form BatchNorm(+ReLU) backward in fp64, on the same fp32 scale/shift/mean/rstd."""
    import torch
    rs = np.random.RandomState(M + N)
    Y = (rs.normal(size=(M, N)) * rs.uniform(0.3, 2.0, N) + rs.uniform(-1, 1, N)).astype(np.float32)
    dZ = np.zeros((M, ldz), np.float32)
    dZ[:, :N] = rs.normal(size=(M, N)).astype(np.float32)
    gamma = rs.uniform(0.5, 1.5, N).astype(np.float32)
    beta = rs.uniform(-0.3, 0.3, N).astype(np.float32)
    p = ffi.ptr
    Yt, dZt, gt, bt = to_cuda(Y), to_cuda(dZ), to_cuda(gamma), to_cuda(beta)
    stats = torch.cat([Yt.double().sum(0), (Yt.double() ** 2).sum(0)]).contiguous()
    sc = torch.empty(N, device="cuda"); sh = torch.empty(N, device="cuda")
    saved = torch.empty(2 * N, device="cuda")
    mm = torch.zeros(N, device="cuda"); mv = torch.ones(N, device="cuda")
    ffi.call("pn2_bn_train_finalize", N, M, p(stats), p(gt), p(bt), 1e-3, 0.9, 1, p(mm), p(mv), p(sc), p(sh), p(saved))
    Y64 = Y.astype(np.float64)
    mean, var = Y64.mean(0), Y64.var(0)
    np.testing.assert_allclose(saved.cpu().numpy()[:N], mean, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(saved.cpu().numpy()[N:], 1 / np.sqrt(var + 1e-3), rtol=1e-6)
    np.testing.assert_allclose(mm.cpu().numpy(), 0.1 * mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mv.cpu().numpy(), 0.9 + 0.1 * var * M / (M - 1), rtol=1e-5)
    red = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    dY = torch.empty((M, N), device="cuda"); dg = torch.zeros(N, device="cuda"); db = torch.zeros(N, device="cuda")
    ffi.call("pn2_bn_bwd_reduce", M, N, p(dZt), ldz, p(Yt), p(sc), p(sh), p(saved), relu, p(red))
    ffi.call("pn2_bn_bwd_apply", M, N, p(dZt), ldz, p(Yt), p(sc), p(sh), p(saved), p(gt), relu, 1, p(red),
             p(dY), p(dg), p(db))
    scn, shn, sv = sc.cpu().numpy().astype(np.float64), sh.cpu().numpy().astype(np.float64), saved.cpu().numpy().astype(np.float64)
    z = Y64 * scn + shn
    dzh = np.where((z > 0) | (relu == 0), dZ[:, :N].astype(np.float64), 0.0)
    xh = (Y64 - sv[:N]) * sv[N:]
    s1, s2 = dzh.sum(0), (dzh * xh).sum(0)
    exp = gamma * sv[N:] * (dzh - s1 / M - xh * s2 / M)
    np.testing.assert_allclose(red.cpu().numpy(), np.concatenate([s1, s2]), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(dY.cpu().numpy(), exp, atol=2e-5)
    np.testing.assert_allclose(dg.cpu().numpy(), s2, rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(db.cpu().numpy(), s1, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,K,N", [(20000, 128, 9), (30000, 6, 32), (3000, 3, 16), (70000, 64, 9)])
def test_narrow_layers_auto_mode(ffi, M, K, N):
    """The 6-channel input layer and the 9-class head in auto mode (-1): forward / dgrad on the
    tensor cores with zero-filled partial chunks, wgrad on the feature-per-thread kernel."""
    import torch
    rs = np.random.RandomState(M + K + N)
    A = rs.normal(size=(M, K)).astype(np.float32)
    W = (rs.uniform(-1, 1, (K, N)) * np.sqrt(6.0 / (K + N))).astype(np.float32)
    b = rs.uniform(-0.5, 0.5, N).astype(np.float32)
    dY = (rs.normal(size=(M, N)) * 0.1).astype(np.float32)
    p = ffi.ptr
    At, Wt, dYt = to_cuda(A), to_cuda(W), to_cuda(dY)
    Y, st = linear_fwd(ffi, At, Wt, to_cuda(b), None, None, False, -1)
    exp = A.astype(np.float64) @ W.astype(np.float64) + b
    np.testing.assert_allclose(Y.cpu().numpy(), exp, atol=1e-5)
    np.testing.assert_allclose(st.cpu().numpy()[:N], exp.sum(0), rtol=1e-5, atol=1e-5 * M)
    ws = _ws(ffi, K, N)
    dX = torch.empty((M, K), dtype=torch.float32, device="cuda")
    ffi.call("pn2_linear_dgrad", M, K, N, p(dYt), p(Wt), p(dX), K, p(ws), ws.numel() * 4, -1)
    np.testing.assert_allclose(dX.cpu().numpy(), dY.astype(np.float64) @ W.astype(np.float64).T, atol=1e-5)
    dW = torch.zeros((K, N), dtype=torch.float32, device="cuda")
    db = torch.zeros(N, dtype=torch.float32, device="cuda")
    ffi.call("pn2_linear_wgrad", M, K, N, p(At), K, None, None, 0, p(dYt), p(dW), p(db), -1)
    expw = A.astype(np.float64).T @ dY.astype(np.float64)
    tol = 2e-5 * max(1.0, np.abs(expw).max())
    np.testing.assert_allclose(dW.cpu().numpy(), expw, atol=tol)
    np.testing.assert_allclose(db.cpu().numpy(), dY.astype(np.float64).sum(0), atol=tol)


@pytest.mark.parametrize("M,K,N,rank4", [(16384, 128, 128, 1), (40000, 67, 64, 0), (4096, 259, 256, 1), (100, 32, 32, 1),
                                        (33000, 32, 9, 0)])
def test_linear_fwd_with_fused_bn_finalize(ffi, M, K, N, rank4):
    """pn2_linear_fwd_bn (the GEMM's last CTA finalises the train-mode BatchNorm) against pn2_linear_fwd +
    pn2_bn_train_finalize: same Y, scale / shift / saved mean+rstd and moving statistics; (100, 32, 32) takes
    the fp32-kernel fallback inside the same entry point."""
    import ctypes
    import torch
    from pn2_b200.util.tf_util import BnFinalize
    rs = np.random.RandomState(M + N)
    lda = (K + 3) // 4 * 4
    A = np.full((M, lda), np.nan, np.float32)
    A[:, :K] = rs.normal(size=(M, K))
    W = (rs.uniform(-1, 1, (K, N)) * np.sqrt(6.0 / (K + N))).astype(np.float32)
    b = rs.uniform(-0.5, 0.5, N).astype(np.float32)
    gamma, beta = rs.uniform(0.5, 1.5, N).astype(np.float32), rs.uniform(-0.3, 0.3, N).astype(np.float32)
    mm0, mv0 = rs.normal(size=N).astype(np.float32), rs.uniform(0.5, 2, N).astype(np.float32)
    p = ffi.ptr
    At, Wt, bt, gt, bet = to_cuda(A), to_cuda(W), to_cuda(b), to_cuda(gamma), to_cuda(beta)
    ws = _ws(ffi, K, N)
    f32 = lambda n: torch.empty(n, dtype=torch.float32, device="cuda")  # noqa: E731
    out = {}
    for fused in (0, 1):
        Y, stats = torch.empty((M, N), dtype=torch.float32, device="cuda"), torch.zeros(2 * N, dtype=torch.float64, device="cuda")
        mm, mv, sc, sh, saved = to_cuda(mm0), to_cuda(mv0), f32(N), f32(N), f32(2 * N)
        if fused:
            counter = torch.zeros(2, dtype=torch.int32, device="cuda")
            fin = BnFinalize(gt.data_ptr(), bet.data_ptr(), mm.data_ptr(), mv.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                             saved.data_ptr(), counter.data_ptr(), 1e-3, 0.7, rank4)
            ffi.call("pn2_linear_fwd_bn", M, K, N, p(At), lda, None, None, 0, p(Wt), p(bt), p(Y), p(stats),
                     ctypes.byref(fin), p(ws), ws.numel() * 4, -1)
        else:
            ffi.call("pn2_linear_fwd", M, K, N, p(At), lda, None, None, 0, p(Wt), p(bt), p(Y), p(stats), p(ws),
                     ws.numel() * 4, -1)
            ffi.call("pn2_bn_train_finalize", N, M, p(stats), p(gt), p(bet), 1e-3, 0.7, rank4, p(mm), p(mv), p(sc),
                     p(sh), p(saved))
        torch.cuda.synchronize()
        out[fused] = [t.cpu().numpy() for t in (Y, sc, sh, saved, mm, mv)]
    for a, b2, name in zip(out[0], out[1], ("Y", "scale", "shift", "saved", "moving_mean", "moving_var")):
        np.testing.assert_allclose(b2, a, rtol=2e-6, atol=1e-7, err_msg=name)
    # and against fp64
    y64 = A[:, :K].astype(np.float64) @ W.astype(np.float64) + b
    mean, var = y64.mean(0), y64.var(0)
    np.testing.assert_allclose(out[1][3][:N], mean, atol=1e-5)
    np.testing.assert_allclose(out[1][1], gamma / np.sqrt(var + 1e-3), rtol=1e-5)
