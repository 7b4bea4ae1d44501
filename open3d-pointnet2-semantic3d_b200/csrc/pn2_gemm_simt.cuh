// pn2_gemm_simt.cuh -- fp32 CUDA-core GEMM with fused prologue / epilogue (sm_100a).
//
// C[M',N'] (=|+=) sum_k f(A)(m,k) * B(k,n) (+ bias[n]),   optional column statistics of C.
//
// One kernel serves the three contractions of a 1x1-conv layer (tf_util.py:128-204):
//   forward  Y  = f(X) W + b      A = X  [M,K]  K-contiguous,  B = W  [K,N] N-contiguous
//   dgrad    dX = dY W^T          A = dY [M,N]  K'-contiguous, B = W  read K'-contiguous
//   wgrad    dW += f(X)^T dY      A = X  read M'-contiguous,   B = dY [M,N] N-contiguous
// f applies the previous layer's BatchNorm affine + ReLU on the fly, indexed by A's
// contiguous coordinate (the feature index in all three cases).
//
// This is the exact-fp32 path (FFMA, round-to-nearest) used for small/ragged shapes, as the
// numerical reference for the tcgen05 3xTF32 kernel, and wherever N' or K' is tiny.
#pragma once
#include "pn2_common.cuh"

namespace pn2 {

constexpr int G_BK = 16;
constexpr int G_THREADS = 256;

template <int T>
struct RowMap {  // thread-tile coordinate i (0..T-1) of thread t (0..15) inside a tile of width B
    template <int B>
    __device__ static __forceinline__ int at(int t, int i) {
        if (T == 8) return (i < 4) ? t * 4 + i : B / 2 + t * 4 + (i - 4);
        return t * T + i;
    }
};

template <int BM, int BN, bool A_KC, bool B_NC, bool ATOMIC>
__global__ void __launch_bounds__(G_THREADS)
gemm_simt_kernel(int Mp, int Np, long Kp, const float *__restrict__ A, long a_sm, long a_sk,
                 const float *__restrict__ B, long b_sk, long b_sn,
                 const float *__restrict__ a_scale, const float *__restrict__ a_shift, int a_relu,
                 const float *__restrict__ bias, float *__restrict__ C, long ldc,
                 double *__restrict__ stats, long k_chunk) {
    pdl_enter();
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int A_PER = BM * G_BK / G_THREADS, B_PER = BN * G_BK / G_THREADS;
    __shared__ __align__(16) float As[G_BK * LDA];
    __shared__ __align__(16) float Bs[G_BK * LDB];
    __shared__ double red[8 * BN], red2[8 * BN];  // column-statistics staging

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const long kbeg = (long)blockIdx.z * k_chunk;
    const long kend = min(Kp, kbeg + k_chunk);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[A_PER], rb[B_PER];

    auto load_tiles = [&](long k0) {
#pragma unroll
        for (int p = 0; p < A_PER; ++p) {
            int mi, ki;
            if (A_KC) {
                ki = tid & 15;
                mi = (tid >> 4) + 16 * p;
            } else {
                mi = tid % BM;
                ki = tid / BM + (G_THREADS / BM) * p;
            }
            const int m = m0 + mi;
            const long k = k0 + ki;
            float v = 0.f;
            if (m < Mp && k < kend) {
                v = __ldg(A + m * a_sm + k * a_sk);
                if (a_scale) {
                    const long f = A_KC ? k : (long)m;
                    v = __fmaf_rn(v, __ldg(a_scale + f), __ldg(a_shift + f));
                    if (a_relu) v = fmaxf(v, 0.f);
                }
            }
            ra[p] = v;
        }
#pragma unroll
        for (int p = 0; p < B_PER; ++p) {
            int ni, ki;
            if (B_NC) {
                ni = tid % BN;
                ki = tid / BN + (G_THREADS / BN) * p;
            } else {
                ki = tid & 15;
                ni = (tid >> 4) + 16 * p;
            }
            const int n = n0 + ni;
            const long k = k0 + ki;
            rb[p] = (n < Np && k < kend) ? __ldg(B + k * b_sk + n * b_sn) : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int p = 0; p < A_PER; ++p) {
            int mi, ki;
            if (A_KC) {
                ki = tid & 15;
                mi = (tid >> 4) + 16 * p;
            } else {
                mi = tid % BM;
                ki = tid / BM + (G_THREADS / BM) * p;
            }
            As[ki * LDA + mi] = ra[p];
        }
#pragma unroll
        for (int p = 0; p < B_PER; ++p) {
            int ni, ki;
            if (B_NC) {
                ni = tid % BN;
                ki = tid / BN + (G_THREADS / BN) * p;
            } else {
                ki = tid & 15;
                ni = (tid >> 4) + 16 * p;
            }
            Bs[ki * LDB + ni] = rb[p];
        }
    };

    if (kbeg < kend) {
        load_tiles(kbeg);
        for (long k0 = kbeg; k0 < kend; k0 += G_BK) {
            store_tiles();
            __syncthreads();
            if (k0 + G_BK < kend) load_tiles(k0 + G_BK);  // register prefetch of the next tile
#pragma unroll
            for (int kk = 0; kk < G_BK; ++kk) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[kk * LDA + RowMap<TM>::template at<BM>(ty, i)];
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[kk * LDB + RowMap<TN>::template at<BN>(tx, j)];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue -------------------------------------------------------------------------
    // column statistics in fp64: E[y^2]-E[y]^2 must not lose the variance to cancellation
    double csum[TN], csq[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) csum[j] = csq[j] = 0.0;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + RowMap<TM>::template at<BM>(ty, i);
        if (m >= Mp) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + RowMap<TN>::template at<BN>(tx, j);
            if (n >= Np) continue;
            float v = acc[i][j];
            if (bias && blockIdx.z == 0) v += __ldg(bias + n);
            if (ATOMIC) atomicAdd(C + m * ldc + n, v);
            else C[m * ldc + n] = v;
            if (stats) {
                csum[j] += (double)v;
                csq[j] = fma((double)v, (double)v, csq[j]);
            }
        }
    }
    if (stats) {
        // lanes l and l^16 share tx; then 8 warps are combined through shared memory
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            csum[j] += __shfl_xor_sync(0xFFFFFFFFu, csum[j], 16);
            csq[j] += __shfl_xor_sync(0xFFFFFFFFu, csq[j], 16);
        }
        const int warp = tid >> 5, lane = tid & 31;
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = RowMap<TN>::template at<BN>(tx, j);
                red[warp * BN + c] = csum[j];
                red2[warp * BN + c] = csq[j];
            }
        }
        __syncthreads();
        for (int c = tid; c < BN; c += G_THREADS) {
            const int n = n0 + c;
            if (n >= Np) continue;
            double s = 0.0, q = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                s += red[w * BN + c];
                q += red2[w * BN + c];
            }
            atomicAdd(stats + n, s);
            atomicAdd(stats + Np + n, q);
        }
    }
}

}  // namespace pn2
