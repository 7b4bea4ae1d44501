// pn2_wgrad_rt.cuh -- register-tiled wgrad for narrow layers (K*N small, M huge), sm_100a.
//
//   dW[K,N] += f(X)[M,K]^T * dY[M,N]
//
// For the first SA layers the output is tiny (6x32 ... 64x128) and the contraction runs over
// 10^5..10^6 rows: tile GEMMs (SIMT or tensor core) waste their tiles on padding and are bound by
// shared-memory traffic.  Here every thread owns a TK x 4 block of dW in registers and streams rows
// straight from global memory (X row fragment + dY row fragment per row: 2-3 vector loads for
// TK*4 FMAs); a group of ceil(K/TK)*ceil(N/4) threads covers dW once and a CTA holds several such
// groups working on interleaved rows.  Partial tiles are combined through shared-memory atomics,
// then one global RED per element per CTA.  Exact fp32 FMA arithmetic.
#pragma once
#include "pn2_common.cuh"

namespace pn2 {

constexpr int WRT_THREADS = 256;
// rows in flight per thread: the kernel is bound by bytes in flight, so 32/TK rows are loaded
// before the first FMA (8 rows for TK=4, 4 rows for TK=8)

template <int TK>
__global__ void __launch_bounds__(WRT_THREADS)
wgrad_rt_kernel(long M, int K, int N, const float *__restrict__ A, int lda,
                const float *__restrict__ a_scale, const float *__restrict__ a_shift, int a_relu,
                const float *__restrict__ dY, float *__restrict__ dW, long rows_per_cta, int tk, int tn) {
    pdl_enter();
    constexpr int WRT_UNROLL = 32 / TK;
    extern __shared__ float tile[];  // [tk*TK][tn*4]
    const int G = tk * tn;
    const int RG = WRT_THREADS / G;
    const int t = threadIdx.x;
    const int rg = t / G, gi = t - rg * G;
    const int ik = gi / tn, in = gi - ik * tn;
    const int k0 = ik * TK, n0 = in * 4;
    const int Np = tn * 4;
    for (int e = t; e < tk * TK * Np; e += WRT_THREADS) tile[e] = 0.f;
    __syncthreads();

    float acc[TK][4];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (rg < RG) {
        float sc[TK], sh[TK];
#pragma unroll
        for (int i = 0; i < TK; ++i) {
            const bool ok = a_scale && (k0 + i < K);
            sc[i] = ok ? __ldg(a_scale + k0 + i) : 1.f;
            sh[i] = ok ? __ldg(a_shift + k0 + i) : 0.f;
        }
        const bool x_vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (k0 + TK <= K);
        const bool y_vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(dY) & 15) == 0) && (n0 + 4 <= N);
        const long r0 = (long)blockIdx.x * rows_per_cta;
        const long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
        for (long m = r0 + rg; m < r1; m += (long)RG * WRT_UNROLL) {
            float x[WRT_UNROLL][TK], y[WRT_UNROLL][4];
#pragma unroll
            for (int u = 0; u < WRT_UNROLL; ++u) {
                const long mm = m + (long)u * RG;
                const bool row_ok = mm < r1;
                const float *xp = A + mm * lda + k0;
                const float *yp = dY + mm * N + n0;
                if (row_ok && x_vec) {
#pragma unroll
                    for (int q = 0; q < TK / 4; ++q) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(xp) + q);
                        x[u][q * 4] = v.x; x[u][q * 4 + 1] = v.y; x[u][q * 4 + 2] = v.z; x[u][q * 4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TK; ++i) x[u][i] = (row_ok && k0 + i < K) ? __ldg(xp + i) : 0.f;
                }
                if (row_ok && y_vec) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(yp));
                    y[u][0] = v.x; y[u][1] = v.y; y[u][2] = v.z; y[u][3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) y[u][j] = (row_ok && n0 + j < N) ? __ldg(yp + j) : 0.f;
                }
                if (a_scale) {
#pragma unroll
                    for (int i = 0; i < TK; ++i) {
                        float v = __fmaf_rn(x[u][i], sc[i], sh[i]);
                        if (a_relu) v = fmaxf(v, 0.f);
                        x[u][i] = (row_ok && k0 + i < K) ? v : 0.f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < WRT_UNROLL; ++u)
#pragma unroll
                for (int i = 0; i < TK; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(x[u][i], y[u][j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TK; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&tile[(k0 + i) * Np + n0 + j], acc[i][j]);
    }
    __syncthreads();
    for (int e = t; e < tk * TK * Np; e += WRT_THREADS) {
        const int k = e / Np, n = e - k * Np;
        if (k < K && n < N) atomicAdd(dW + (long)k * N + n, tile[e]);
    }
}

// returns PN2_EUNSUPPORTED when the shape does not fit one thread group per CTA
static int wgrad_rt(long M, int K, int N, const float *A, int lda, const float *a_scale,
                    const float *a_shift, int a_relu, const float *dY, float *dW, cudaStream_t st) {
    const int tn = ceil_div(N, 4);
    int TKsel = 4, tk = ceil_div(K, 4);
    if (tk * tn > WRT_THREADS) {
        TKsel = 8;
        tk = ceil_div(K, 8);
    }
    if (tk * tn > WRT_THREADS || M < 4096) return PN2_EUNSUPPORTED;
    const int RG = WRT_THREADS / (tk * tn);
    long ctas = ceil_div<long>(M, (long)RG * 256);  // >= 256 rows per row-group
    const long cap = (long)num_sms() * 4;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    const long rows_per_cta = ceil_div<long>(M, ctas);
    const size_t smem = (size_t)tk * TKsel * tn * 4 * sizeof(float);
    if (TKsel == 4)
        launch_k(wgrad_rt_kernel<4>, (unsigned)ctas, WRT_THREADS, smem, st, M, K, N, A, lda, a_scale, a_shift, a_relu,
                                                                     dY, dW, rows_per_cta, tk, tn);
    else
        launch_k(wgrad_rt_kernel<8>, (unsigned)ctas, WRT_THREADS, smem, st, M, K, N, A, lda, a_scale, a_shift, a_relu,
                                                                     dY, dW, rows_per_cta, tk, tn);
    return finish_launch();
}

}  // namespace pn2
