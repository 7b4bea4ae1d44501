// pn2_dense.cu -- shared-MLP layer pieces for sm_100a: linear fwd/dgrad/wgrad dispatch,
// train/eval BatchNorm, fused affine+ReLU(+max-pool), their backward passes, dropout, the
// weighted softmax cross-entropy and the Adam update.
//
// These replace the TensorFlow graph ops the reference strings between its custom ops
// (util/tf_util.py:54-204, 555-581, 646-665; util/pointnet_util.py:150-170, 313-324;
// model.py:132-161; train.py:387-388).  Nothing here has a reference kernel to follow.
#include <math.h>

#include "pn2_common.cuh"
#include "pn2_gemm_simt.cuh"
#include "pn2_wgrad_rt.cuh"

namespace pn2 {

// implemented in pn2_gemm_tc.cu (tcgen05 3xTF32 path); returns PN2_EUNSUPPORTED for shapes it
// does not cover so that the caller can fall back to the exact fp32 kernel.
int tc_linear_fwd(long M, int K, int N, const float *A, int lda, const float *a_scale,
                  const float *a_shift, int a_relu, const float *W, const float *bias, float *Y,
                  double *stats, float *ws, size_t ws_bytes, bool image_ready, const pn2_bn_finalize *fin,
                  cudaStream_t st);
int tc_linear_dgrad(long M, int K, int N, const float *dY, const float *W, float *dX, int ldx,
                    float *ws, size_t ws_bytes, bool image_ready, cudaStream_t st);
int tc_describe_image(int K, int N, bool dgrad, const float *W, float *image, pn2_linear_image *out);
size_t tc_image_bytes_for(int K, int N, bool dgrad);
int tc_prepare_images(int count, const pn2_linear_image *table_dev, cudaStream_t st);
size_t tc_workspace_bytes(int K, int N);
int tc_linear_wgrad(long M, int K, int N, const float *A, int lda, const float *a_scale,
                    const float *a_shift, int a_relu, const float *dY, float *dW, bool force,
                    int *k_done, cudaStream_t st);

// ---- SIMT GEMM dispatch ----------------------------------------------------------------------
template <bool A_KC, bool B_NC, bool ATOMIC>
static int launch_gemm(int Mp, int Np, long Kp, const float *A, long a_sm, long a_sk,
                       const float *B, long b_sk, long b_sn, const float *a_scale,
                       const float *a_shift, int a_relu, const float *bias, float *C, long ldc,
                       double *stats, int splits, cudaStream_t st) {
    const int bm = Mp > 64 ? 128 : (Mp > 32 ? 64 : 32);
    const int bn = Np > 64 ? 128 : (Np > 32 ? 64 : 32);
    long k_chunk = ceil_div<long>(Kp, splits);
    k_chunk = ceil_div<long>(k_chunk, G_BK) * G_BK;
    splits = (int)ceil_div<long>(Kp, k_chunk);
    if (splits < 1) splits = 1;
    dim3 grid((unsigned)ceil_div(Mp, bm), (unsigned)ceil_div(Np, bn), (unsigned)splits);
#define PN2_G(BM_, BN_)                                                                        \
    launch_k(gemm_simt_kernel<BM_, BN_, A_KC, B_NC, ATOMIC>, grid, G_THREADS, 0, st,                  \
        Mp, Np, Kp, A, a_sm, a_sk, B, b_sk, b_sn, a_scale, a_shift, a_relu, bias, C, ldc, stats, \
        k_chunk)
    if (bm == 128 && bn == 128) PN2_G(128, 128);
    else if (bm == 128 && bn == 64) PN2_G(128, 64);
    else if (bm == 128 && bn == 32) PN2_G(128, 32);
    else if (bm == 64 && bn == 128) PN2_G(64, 128);
    else if (bm == 64 && bn == 64) PN2_G(64, 64);
    else if (bm == 64 && bn == 32) PN2_G(64, 32);
    else if (bm == 32 && bn == 128) PN2_G(32, 128);
    else if (bm == 32 && bn == 64) PN2_G(32, 64);
    else PN2_G(32, 32);
#undef PN2_G
    return finish_launch();
}

// ---- column-wise helpers over a row-major [M,N] matrix ------------------------------------------
// Block = 256 threads owning a slab of rows; thread t works on columns c = cx, cx+lanes, ... and
// rows r = ry, ry+rpp, ... so that a warp reads contiguous columns of one (or a few) rows.
// per-block combination of fp64 column sums: shared-memory atomics, then ONE global atomic per
// column per block (thousands of threads hammering 2N global addresses serialise in L2)
__device__ __forceinline__ void block_zero2(double *sred, int N) {
    for (int e = threadIdx.x; e < 2 * N; e += blockDim.x) sred[e] = 0.0;
    __syncthreads();
}
__device__ __forceinline__ void block_add2(double *sred, int N, int c, double a0, double a1) {
    atomicAdd(sred + c, a0);
    atomicAdd(sred + N + c, a1);
}
__device__ __forceinline__ void block_flush2(double *sred, int N, double *__restrict__ red) {
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * N; e += blockDim.x) atomicAdd(red + e, sred[e]);
}

struct Slab {
    int lanes, rpp, cx, ry;
    long r0, r1;
    bool active;
};
__device__ __forceinline__ Slab make_slab(long M, int N, long rows_per_block) {
    Slab s;
    s.lanes = N < (int)blockDim.x ? N : (int)blockDim.x;
    s.rpp = blockDim.x / s.lanes;
    s.cx = threadIdx.x % s.lanes;
    s.ry = threadIdx.x / s.lanes;
    s.active = s.ry < s.rpp;
    s.r0 = (long)blockIdx.x * rows_per_block;
    s.r1 = s.r0 + rows_per_block < M ? s.r0 + rows_per_block : M;
    return s;
}

// dZh, xhat for one element
__device__ __forceinline__ void bn_elem(float y, float dz, float sc, float sh, float mean,
                                        float rstd, int relu, float &dzh, float &xhat) {
    const float z = __fmaf_rn(y, sc, sh);
    dzh = (relu && !(z > 0.f)) ? 0.f : dz;
    xhat = (y - mean) * rstd;
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(long M, int N, long rpb, const float *__restrict__ dZ, int ldz,
                     const float *__restrict__ Y, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ saved, int relu,
                     double *__restrict__ red) {
    pdl_enter();
    extern __shared__ double sred[];  // [2N] per-block partial sums
    const Slab s = make_slab(M, N, rpb);
    block_zero2(sred, N);
    for (int c = s.cx; s.active && c < N; c += s.lanes) {
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float mean = __ldg(saved + c), rstd = __ldg(saved + N + c);
        double a0 = 0.0, a1 = 0.0;
        for (long r = s.r0 + s.ry; r < s.r1; r += s.rpp) {
            float dzh, xh;
            bn_elem(__ldg(Y + r * N + c), __ldg(dZ + r * ldz + c), sc, sh, mean, rstd, relu, dzh, xh);
            a0 += (double)dzh;
            a1 = fma((double)dzh, (double)xh, a1);
        }
        block_add2(sred, N, c, a0, a1);
    }
    block_flush2(sred, N, red);
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(long M, int N, long rpb, const float *__restrict__ dZ, int ldz,
                    const float *__restrict__ Y, const float *__restrict__ scale,
                    const float *__restrict__ shift, const float *__restrict__ saved,
                    const float *__restrict__ gamma, int relu, int bn,
                    const double *__restrict__ red, float *__restrict__ dY,
                    float *__restrict__ dgamma, float *__restrict__ dbeta) {
    pdl_enter();
    const Slab s = make_slab(M, N, rpb);
    if (!s.active) return;
    const double invM = 1.0 / (double)M;
    for (int c = s.cx; c < N; c += s.lanes) {
        float sc = 1.f, sh = 0.f, mean = 0.f, rstd = 1.f, gs = 1.f, m0 = 0.f, m1 = 0.f;
        if (scale) {
            sc = __ldg(scale + c);
            sh = __ldg(shift + c);
        }
        if (bn) {
            mean = __ldg(saved + c);
            rstd = __ldg(saved + N + c);
            gs = __ldg(gamma + c) * rstd;
            m0 = (float)(red[c] * invM);
            m1 = (float)(red[N + c] * invM);
            if (blockIdx.x == 0 && s.ry == 0) {
                if (dgamma) dgamma[c] += (float)red[N + c];
                if (dbeta) dbeta[c] += (float)red[c];
            }
        }
        for (long r = s.r0 + s.ry; r < s.r1; r += s.rpp) {
            float dzh, xh;
            bn_elem(__ldg(Y + r * N + c), __ldg(dZ + r * ldz + c), sc, sh, mean, rstd, relu, dzh, xh);
            dY[r * N + c] = bn ? gs * (dzh - m0 - xh * m1) : dzh;
        }
    }
}

// ---- float4-vectorised variants (N % 4 == 0, ldz % 4 == 0, 16-byte aligned bases) -----------------
// A block owns a slab of rows; thread t works on the column quad cq = t % (N/4) and rows
// ry, ry+rpp, ...; four rows are in flight per thread (8 independent 16-byte loads).
struct Slab4 {
    int nq, rpp, cq, ry;
    long r0, r1;
    bool active;
};
__device__ __forceinline__ Slab4 make_slab4(long M, int N, long rows_per_block) {
    Slab4 s;
    const int q = N >> 2;
    s.nq = q < (int)blockDim.x ? q : (int)blockDim.x;
    s.rpp = blockDim.x / s.nq;
    s.cq = threadIdx.x % s.nq;
    s.ry = threadIdx.x / s.nq;
    s.active = s.ry < s.rpp;
    s.r0 = (long)blockIdx.x * rows_per_block;
    s.r1 = s.r0 + rows_per_block < M ? s.r0 + rows_per_block : M;
    return s;
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_v4_kernel(long M, int N, long rpb, const float *__restrict__ dZ, int ldz,
                        const float *__restrict__ Y, const float *__restrict__ scale,
                        const float *__restrict__ shift, const float *__restrict__ saved, int relu,
                        double *__restrict__ red) {
    pdl_enter();
    extern __shared__ double sred[];  // [2N] per-block partial sums
    const Slab4 s = make_slab4(M, N, rpb);
    block_zero2(sred, N);
    const int nq = N >> 2;
    for (int cq = s.cq; s.active && cq < nq; cq += s.nq) {
        const int c = cq * 4;
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
        const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
        const float4 mean = __ldg(reinterpret_cast<const float4 *>(saved + c));
        const float4 rstd = __ldg(reinterpret_cast<const float4 *>(saved + N + c));
        double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
        for (long r = s.r0 + s.ry; r < s.r1; r += 4L * s.rpp) {
            float4 y[4], z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long rr = r + (long)u * s.rpp;
                if (rr < s.r1) {
                    y[u] = __ldg(reinterpret_cast<const float4 *>(Y + rr * N + c));
                    z[u] = __ldg(reinterpret_cast<const float4 *>(dZ + rr * ldz + c));
                } else {
                    y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    z[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // fp32 partial sums over the four rows in flight, fp64 across batches
            float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float dzh, xh;
                bn_elem(y[u].x, z[u].x, sc.x, sh.x, mean.x, rstd.x, relu, dzh, xh); f0[0] += dzh; f1[0] = __fmaf_rn(dzh, xh, f1[0]);
                bn_elem(y[u].y, z[u].y, sc.y, sh.y, mean.y, rstd.y, relu, dzh, xh); f0[1] += dzh; f1[1] = __fmaf_rn(dzh, xh, f1[1]);
                bn_elem(y[u].z, z[u].z, sc.z, sh.z, mean.z, rstd.z, relu, dzh, xh); f0[2] += dzh; f1[2] = __fmaf_rn(dzh, xh, f1[2]);
                bn_elem(y[u].w, z[u].w, sc.w, sh.w, mean.w, rstd.w, relu, dzh, xh); f0[3] += dzh; f1[3] = __fmaf_rn(dzh, xh, f1[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0[j] += (double)f0[j];
                a1[j] += (double)f1[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) block_add2(sred, N, c + j, a0[j], a1[j]);
    }
    block_flush2(sred, N, red);
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_v4_kernel(long M, int N, long rpb, const float *__restrict__ dZ, int ldz,
                       const float *__restrict__ Y, const float *__restrict__ scale,
                       const float *__restrict__ shift, const float *__restrict__ saved,
                       const float *__restrict__ gamma, int relu, int bn,
                       const double *__restrict__ red, float *__restrict__ dY,
                       float *__restrict__ dgamma, float *__restrict__ dbeta) {
    pdl_enter();
    const Slab4 s = make_slab4(M, N, rpb);
    if (!s.active) return;
    const double invM = 1.0 / (double)M;
    const int nq = N >> 2;
    for (int cq = s.cq; cq < nq; cq += s.nq) {
        const int c = cq * 4;
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, mean[4] = {0.f, 0.f, 0.f, 0.f},
              rstd[4] = {1.f, 1.f, 1.f, 1.f}, gs[4] = {1.f, 1.f, 1.f, 1.f}, m0[4] = {0.f, 0.f, 0.f, 0.f},
              m1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (scale) {
                sc[j] = __ldg(scale + c + j);
                sh[j] = __ldg(shift + c + j);
            }
            if (bn) {
                mean[j] = __ldg(saved + c + j);
                rstd[j] = __ldg(saved + N + c + j);
                gs[j] = __ldg(gamma + c + j) * rstd[j];
                m0[j] = (float)(red[c + j] * invM);
                m1[j] = (float)(red[N + c + j] * invM);
                if (blockIdx.x == 0 && s.ry == 0) {
                    if (dgamma) dgamma[c + j] += (float)red[N + c + j];
                    if (dbeta) dbeta[c + j] += (float)red[c + j];
                }
            }
        }
        for (long r = s.r0 + s.ry; r < s.r1; r += 4L * s.rpp) {
            float4 y[4], z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long rr = r + (long)u * s.rpp;
                if (rr < s.r1) {
                    y[u] = __ldg(reinterpret_cast<const float4 *>(Y + rr * N + c));
                    z[u] = __ldg(reinterpret_cast<const float4 *>(dZ + rr * ldz + c));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long rr = r + (long)u * s.rpp;
                if (rr < s.r1) {
                    const float *yy = reinterpret_cast<const float *>(&y[u]);
                    const float *zz = reinterpret_cast<const float *>(&z[u]);
                    float4 o;
                    float *oo = reinterpret_cast<float *>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float dzh, xh;
                        bn_elem(yy[j], zz[j], sc[j], sh[j], mean[j], rstd[j], relu, dzh, xh);
                        oo[j] = bn ? gs[j] * (dzh - m0[j] - xh * m1[j]) : dzh;
                    }
                    *reinterpret_cast<float4 *>(dY + rr * N + c) = o;
                }
            }
        }
    }
}

// pooled upstream gradient: dZh[g*ns+j, c] = dOut[g,c] if arg[g,c]==j (and z>0) else 0
__global__ void __launch_bounds__(256)
bn_bwd_reduce_pool_kernel(long G, int ns, int N, const float *__restrict__ dOut,
                          const int *__restrict__ arg, const float *__restrict__ Y,
                          const float *__restrict__ scale, const float *__restrict__ shift,
                          const float *__restrict__ saved, int relu, double *__restrict__ red) {
    pdl_enter();
    extern __shared__ double sred[];  // [2N] per-block partial sums
    const long rpb = ceil_div<long>(G, (long)gridDim.x);
    const Slab s = make_slab(G, N, rpb);
    block_zero2(sred, N);
    for (int c = s.cx; s.active && c < N; c += s.lanes) {
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float mean = __ldg(saved + c), rstd = __ldg(saved + N + c);
        double a0 = 0.0, a1 = 0.0;
        for (long g = s.r0 + s.ry; g < s.r1; g += s.rpp) {
            const int j = __ldg(arg + g * N + c);
            float dzh, xh;
            bn_elem(__ldg(Y + (g * ns + j) * N + c), __ldg(dOut + g * N + c), sc, sh, mean, rstd,
                    relu, dzh, xh);
            a0 += (double)dzh;
            a1 = fma((double)dzh, (double)xh, a1);
        }
        block_add2(sred, N, c, a0, a1);
    }
    block_flush2(sred, N, red);
}

// float4 variant of the kernel below (N % 4 == 0, 16-byte aligned bases): a block owns a slab of
// groups; thread t works on column quad cq = t % (N/4) and sample rows j = ry, ry+rpp, ...; the
// group's arg / dOut quads are loaded once per group
__global__ void __launch_bounds__(256)
bn_bwd_apply_pool_v4_kernel(long G, int ns, int N, long gpb, const float *__restrict__ dOut,
                            const int *__restrict__ arg, const float *__restrict__ Y,
                            const float *__restrict__ scale, const float *__restrict__ shift,
                            const float *__restrict__ saved, const float *__restrict__ gamma,
                            int relu, int bn, const double *__restrict__ red,
                            float *__restrict__ dY, float *__restrict__ dgamma,
                            float *__restrict__ dbeta) {
    pdl_enter();
    const Slab4 s = make_slab4(G, N, gpb);
    if (!s.active) return;
    const double invM = 1.0 / (double)(G * ns);
    const int nq = N >> 2;
    for (int cq = s.cq; cq < nq; cq += s.nq) {
        const int c = cq * 4;
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, mean[4] = {0.f, 0.f, 0.f, 0.f},
              rstd[4] = {1.f, 1.f, 1.f, 1.f}, gs[4] = {1.f, 1.f, 1.f, 1.f}, m0[4] = {0.f, 0.f, 0.f, 0.f},
              m1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (scale) {
                sc[j] = __ldg(scale + c + j);
                sh[j] = __ldg(shift + c + j);
            }
            if (bn) {
                mean[j] = __ldg(saved + c + j);
                rstd[j] = __ldg(saved + N + c + j);
                gs[j] = __ldg(gamma + c + j) * rstd[j];
                m0[j] = (float)(red[c + j] * invM);
                m1[j] = (float)(red[N + c + j] * invM);
                if (blockIdx.x == 0 && s.ry == 0) {
                    if (dgamma) dgamma[c + j] += (float)red[N + c + j];
                    if (dbeta) dbeta[c + j] += (float)red[c + j];
                }
            }
        }
        for (long g = s.r0; g < s.r1; ++g) {
            const int4 a4 = __ldg(reinterpret_cast<const int4 *>(arg + g * N + c));
            const float4 d4 = __ldg(reinterpret_cast<const float4 *>(dOut + g * N + c));
            const int aa[4] = {a4.x, a4.y, a4.z, a4.w};
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            const float *yg = Y + g * ns * N + c;
            float *og = dY + g * ns * N + c;
#pragma unroll 4
            for (int j = s.ry; j < ns; j += s.rpp) {
                const float4 y4 = __ldg(reinterpret_cast<const float4 *>(yg + (long)j * N));
                const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
                float4 o;
                float *oo = reinterpret_cast<float *>(&o);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float dzh, xh;
                    bn_elem(yy[q], aa[q] == j ? dd[q] : 0.f, sc[q], sh[q], mean[q], rstd[q], relu, dzh, xh);
                    oo[q] = bn ? gs[q] * (dzh - m0[q] - xh * m1[q]) : dzh;
                }
                *reinterpret_cast<float4 *>(og + (long)j * N) = o;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_pool_kernel(long G, int ns, int N, long rpb, const float *__restrict__ dOut,
                         const int *__restrict__ arg, const float *__restrict__ Y,
                         const float *__restrict__ scale, const float *__restrict__ shift,
                         const float *__restrict__ saved, const float *__restrict__ gamma,
                         int relu, int bn, const double *__restrict__ red,
                         float *__restrict__ dY, float *__restrict__ dgamma,
                         float *__restrict__ dbeta) {
    pdl_enter();
    const long M = G * ns;
    const Slab s = make_slab(M, N, rpb);
    if (!s.active) return;
    const double invM = 1.0 / (double)M;
    for (int c = s.cx; c < N; c += s.lanes) {
        float sc = 1.f, sh = 0.f, mean = 0.f, rstd = 1.f, gs = 1.f, m0 = 0.f, m1 = 0.f;
        if (scale) {
            sc = __ldg(scale + c);
            sh = __ldg(shift + c);
        }
        if (bn) {
            mean = __ldg(saved + c);
            rstd = __ldg(saved + N + c);
            gs = __ldg(gamma + c) * rstd;
            m0 = (float)(red[c] * invM);
            m1 = (float)(red[N + c] * invM);
            if (blockIdx.x == 0 && s.ry == 0) {
                if (dgamma) dgamma[c] += (float)red[N + c];
                if (dbeta) dbeta[c] += (float)red[c];
            }
        }
        for (long r = s.r0 + s.ry; r < s.r1; r += s.rpp) {
            const long g = r / ns;
            const int j = (int)(r - g * ns);
            const float dz = (__ldg(arg + g * N + c) == j) ? __ldg(dOut + g * N + c) : 0.f;
            float dzh, xh;
            bn_elem(__ldg(Y + r * N + c), dz, sc, sh, mean, rstd, relu, dzh, xh);
            dY[r * N + c] = bn ? gs * (dzh - m0 - xh * m1) : dzh;
        }
    }
}

__global__ void bn_train_finalize_kernel(int N, long M, const double *__restrict__ stats,
                                         const float *__restrict__ gamma,
                                         const float *__restrict__ beta, float eps, float decay,
                                         int unbiased_moving, float *__restrict__ moving_mean,
                                         float *__restrict__ moving_var, float *__restrict__ scale,
                                         float *__restrict__ shift, float *__restrict__ saved) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const double mean = stats[c] / (double)M;
    double var = stats[N + c] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const float sc = (float)((double)gamma[c] * rstd);
    scale[c] = sc;
    shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * rstd);
    saved[c] = (float)mean;
    saved[N + c] = (float)rstd;
    if (moving_mean) {
        const double uv = unbiased_moving && M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        // assign_moving_average: v -= (v - value) * (1 - decay)
        moving_mean[c] = (float)((double)moving_mean[c] - ((double)moving_mean[c] - mean) * (1.0 - (double)decay));
        moving_var[c] = (float)((double)moving_var[c] - ((double)moving_var[c] - uv) * (1.0 - (double)decay));
    }
}

__global__ void bn_eval_affine_kernel(int N, const float *__restrict__ gamma,
                                      const float *__restrict__ beta,
                                      const float *__restrict__ mm, const float *__restrict__ mv,
                                      float eps, float *__restrict__ scale,
                                      float *__restrict__ shift) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const double rstd = 1.0 / sqrt((double)mv[c] + (double)eps);
    scale[c] = (float)((double)gamma[c] * rstd);
    shift[c] = (float)((double)beta[c] - (double)mm[c] * (double)gamma[c] * rstd);
}

__global__ void affine_act_kernel(long total, int N, const float *__restrict__ Y,
                                  const float *__restrict__ scale,
                                  const float *__restrict__ shift, int relu,
                                  float *__restrict__ Z, int ldz) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const long r = e / N;
        const int c = (int)(e - r * N);
        float v = __ldg(Y + e);
        if (scale) v = __fmaf_rn(v, __ldg(scale + c), __ldg(shift + c));
        if (relu) v = fmaxf(v, 0.f);
        Z[r * ldz + c] = v;
    }
}

// float4 variants (N % 4 == 0, ldz % 4 == 0, 16-byte aligned bases)
__global__ void affine_act_v4_kernel(long total4, int nq, const float *__restrict__ Y,
                                     const float *__restrict__ scale,
                                     const float *__restrict__ shift, int relu,
                                     float *__restrict__ Z, int ldz) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4;
         e += (long)gridDim.x * blockDim.x) {
        const long r = e / nq;
        const int c = (int)(e - r * nq) * 4;
        float4 v = __ldg(reinterpret_cast<const float4 *>(Y) + e);
        if (scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
            const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
            v.x = __fmaf_rn(v.x, sc.x, sh.x);
            v.y = __fmaf_rn(v.y, sc.y, sh.y);
            v.z = __fmaf_rn(v.z, sc.z, sh.z);
            v.w = __fmaf_rn(v.w, sc.w, sh.w);
        }
        if (relu) {
            v.x = fmaxf(v.x, 0.f);
            v.y = fmaxf(v.y, 0.f);
            v.z = fmaxf(v.z, 0.f);
            v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4 *>(Z + r * ldz + c) = v;
    }
}

__global__ void __launch_bounds__(128)
affine_act_maxpool_v4_kernel(long total4, int ns, int N, const float *__restrict__ Y,
                             const float *__restrict__ scale, const float *__restrict__ shift,
                             int relu, float *__restrict__ out, int *__restrict__ arg) {
    pdl_enter();
    const int nq = N >> 2;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4;
         e += (long)gridDim.x * blockDim.x) {
        const long g = e / nq;
        const int c = (int)(e - g * nq) * 4;
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (scale) {
            const float4 s4 = __ldg(reinterpret_cast<const float4 *>(scale + c));
            const float4 h4 = __ldg(reinterpret_cast<const float4 *>(shift + c));
            sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
            sh[0] = h4.x; sh[1] = h4.y; sh[2] = h4.z; sh[3] = h4.w;
        }
        const float *y = Y + g * ns * N + c;
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int bj[4] = {0, 0, 0, 0};
#pragma unroll 8
        for (int j = 0; j < ns; ++j) {
            const float4 y4 = __ldg(reinterpret_cast<const float4 *>(y + (long)j * N));
            const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = __fmaf_rn(yy[q], sc[q], sh[q]);
                if (relu) v = fmaxf(v, 0.f);
                if (v > best[q]) {
                    best[q] = v;
                    bj[q] = j;
                }
            }
        }
        *reinterpret_cast<float4 *>(out + g * N + c) = make_float4(best[0], best[1], best[2], best[3]);
        if (arg) *reinterpret_cast<int4 *>(arg + g * N + c) = make_int4(bj[0], bj[1], bj[2], bj[3]);
    }
}

__global__ void affine_act_maxpool_kernel(long total, int ns, int N, const float *__restrict__ Y,
                                          const float *__restrict__ scale,
                                          const float *__restrict__ shift, int relu,
                                          float *__restrict__ out, int *__restrict__ arg) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const long g = e / N;
        const int c = (int)(e - g * N);
        const float sc = scale ? __ldg(scale + c) : 1.f, sh = scale ? __ldg(shift + c) : 0.f;
        const float *y = Y + g * ns * N + c;
        float best = -INFINITY;
        int bj = 0;
#pragma unroll 4
        for (int j = 0; j < ns; ++j) {
            float v = __fmaf_rn(__ldg(y + (long)j * N), sc, sh);
            if (relu) v = fmaxf(v, 0.f);
            if (v > best) {
                best = v;
                bj = j;
            }
        }
        out[e] = best;
        if (arg) arg[e] = bj;
    }
}

// ---- alternative poolings of pointnet_sa_module (pointnet_util.py:171-191) ---------------------------
// X[G*ns, N] = the chain's activated output.  mode 1 "avg": reduce_mean over nsample; mode 2
// "weighted_avg": sum_j X * w_j with w = exp(-5|g|) / sum_j exp(-5|g|) of the centred grouped xyz;
// mode 3 "max_and_avg": [avg | max] (the reference concatenates in that order, :187-191), arg = first j
// attaining the maximum.  One thread per (group, channel); sums run j = 0..ns-1 in fp32.
__global__ void pool_weights_kernel(long G, int ns, const float *__restrict__ gxyz, int ld,
                                    float *__restrict__ w) {
    pdl_enter();
    for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < G; g += (long)gridDim.x * blockDim.x) {
        float tot = 0.f;
        for (int j = 0; j < ns; ++j) {
            const float *p = gxyz + (g * ns + j) * (long)ld;
            const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
            const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
            const float e = expf(-dist * 5.f);
            w[g * ns + j] = e;
            tot += e;
        }
        for (int j = 0; j < ns; ++j) w[g * ns + j] = w[g * ns + j] / tot;
    }
}

__global__ void group_pool_kernel(long total, int ns, int N, const float *__restrict__ X,
                                  const float *__restrict__ w, int mode, float *__restrict__ out,
                                  int *__restrict__ arg) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long g = e / N;
        const int c = (int)(e - g * N);
        const float *x = X + g * ns * (long)N + c;
        float sum = 0.f, mx = -INFINITY;
        int am = 0;
        for (int j = 0; j < ns; ++j) {
            const float v = __ldg(x + (long)j * N);
            sum = __fadd_rn(sum, mode == 2 ? __fmul_rn(v, __ldg(w + g * ns + j)) : v);
            if (v > mx) {
                mx = v;
                am = j;
            }
        }
        if (mode == 1) out[e] = sum / (float)ns;
        else if (mode == 2) out[e] = sum;
        else {
            out[g * 2 * N + c] = sum / (float)ns;
            out[g * 2 * N + N + c] = mx;
            arg[e] = am;
        }
    }
}

__global__ void group_pool_grad_kernel(long total, int ns, int N, const float *__restrict__ dOut,
                                       const float *__restrict__ w, const int *__restrict__ arg, int mode,
                                       float *__restrict__ dX) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long row = e / N;
        const int c = (int)(e - row * N);
        const long g = row / ns;
        const int j = (int)(row - g * ns);
        float v;
        if (mode == 1) v = __ldg(dOut + g * N + c) / (float)ns;
        else if (mode == 2) v = __fmul_rn(__ldg(dOut + g * N + c), __ldg(w + row));
        else
            v = __ldg(dOut + g * 2 * N + c) / (float)ns +
                (__ldg(arg + g * N + c) == j ? __ldg(dOut + g * 2 * N + N + c) : 0.f);
        dX[e] = v;
    }
}

// ---- test hook: the ReLU mask exactly as every kernel of the chain evaluates it --------------------
__global__ void relu_mask_kernel(long total, int N, const float *__restrict__ Y,
                                 const float *__restrict__ scale, const float *__restrict__ shift,
                                 unsigned char *__restrict__ mask) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % N);
        const float y = __ldg(Y + e);
        const float z = scale ? __fmaf_rn(y, __ldg(scale + c), __ldg(shift + c)) : y;
        mask[e] = z > 0.f ? 1 : 0;
    }
}

// ---- dropout: counter-based generator (splitmix64 finaliser over seed ^ index) -------------------
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, long i, float keep_prob) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);  // [0,1)
    return u < keep_prob;
}
__global__ void dropout_kernel(long n, const float *__restrict__ x, float keep_prob, float inv,
                               unsigned long long seed,
                               const unsigned long long *__restrict__ seed_dev,
                               float *__restrict__ out) {
    pdl_enter();
    if (seed_dev) seed += *seed_dev;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
         i += (long)gridDim.x * blockDim.x)
        out[i] = dropout_keep(seed, i, keep_prob) ? __fmul_rn(__ldg(x + i), inv) : 0.f;
}
// four elements per thread (16-byte loads / stores); element i still uses hash(seed, i), so the mask is the same
__global__ void dropout_v4_kernel(long n4, const float4 *__restrict__ x, float keep_prob, float inv,
                                  unsigned long long seed, const unsigned long long *__restrict__ seed_dev,
                                  float4 *__restrict__ out) {
    pdl_enter();
    if (seed_dev) seed += *seed_dev;
    for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < n4; q += (long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + q);
        float4 o;
        o.x = dropout_keep(seed, 4 * q, keep_prob) ? __fmul_rn(v.x, inv) : 0.f;
        o.y = dropout_keep(seed, 4 * q + 1, keep_prob) ? __fmul_rn(v.y, inv) : 0.f;
        o.z = dropout_keep(seed, 4 * q + 2, keep_prob) ? __fmul_rn(v.z, inv) : 0.f;
        o.w = dropout_keep(seed, 4 * q + 3, keep_prob) ? __fmul_rn(v.w, inv) : 0.f;
        out[q] = o;
    }
}
__global__ void dropout_mask_kernel(long n, float keep_prob, unsigned long long seed,
                                    unsigned char *__restrict__ mask) {
    pdl_enter();
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
         i += (long)gridDim.x * blockDim.x)
        mask[i] = dropout_keep(seed, i, keep_prob) ? 1 : 0;
}

// ---- weighted sparse softmax cross entropy ---------------------------------------------------------
__global__ void softmax_ce_reduce_kernel(long rows, int C, const float *__restrict__ logits,
                                         const int *__restrict__ labels,
                                         const float *__restrict__ weights,
                                         double *__restrict__ acc) {
    pdl_enter();
    double s = 0.0, nz = 0.0;
    for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < rows;
         r += (long)gridDim.x * blockDim.x) {
        const float *x = logits + r * C;
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(x + c));
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(__ldg(x + c) - mx);
        const float lse = mx + logf(se);
        const float w = weights ? __ldg(weights + r) : 1.f;
        // a label outside [0,C) must not index the logits: TF's GPU kernel yields NaN for such a row
        // (tf.nn.sparse_softmax_cross_entropy_with_logits), and so does this one -- loudly visible
        const int lab = __ldg(labels + r);
        const float ce = (lab >= 0 && lab < C) ? lse - __ldg(x + lab) : __int_as_float(0x7fc00000);
        s += (double)(ce * w);
        nz += (w != 0.f) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
        nz += __shfl_xor_sync(0xFFFFFFFFu, nz, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, s);
        atomicAdd(acc + 1, nz);
    }
}
__global__ void softmax_ce_grad_kernel(long rows, int C, const float *__restrict__ logits,
                                       const int *__restrict__ labels,
                                       const float *__restrict__ weights,
                                       const double *__restrict__ acc, float gscale,
                                       const float *__restrict__ gscale_dev,
                                       float *__restrict__ loss, float *__restrict__ dlogits) {
    pdl_enter();
    const double nz = acc[1] > 0.0 ? acc[1] : 1.0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) *loss = (float)(acc[0] / nz);
    if (!dlogits) return;
    const float inv = (float)((double)gscale * (gscale_dev ? (double)*gscale_dev : 1.0) / nz);
    for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < rows;
         r += (long)gridDim.x * blockDim.x) {
        const float *x = logits + r * C;
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(x + c));
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(__ldg(x + c) - mx);
        const float w = (weights ? __ldg(weights + r) : 1.f) * inv;
        const int lab = __ldg(labels + r);
        const float rse = 1.f / se;
        const bool lab_ok = lab >= 0 && lab < C;  // out of range: NaN row, like the loss
        for (int c = 0; c < C; ++c) {
            const float p = expf(__ldg(x + c) - mx) * rse;
            dlogits[r * C + c] = lab_ok ? w * (p - (c == lab ? 1.f : 0.f)) : __int_as_float(0x7fc00000);
        }
    }
}

__global__ void adam_kernel(long n, float *__restrict__ p, const float *__restrict__ g,
                            float *__restrict__ m, float *__restrict__ v, float lr_t, float b1,
                            float b2, float eps, float gscale) {
    pdl_enter();
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
         i += (long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// Second generation of the narrow-output wgrad (the 9-class head: M = 131072, K = 128, N = 9 ran at 0.75 TB/s
// with the kernel above -- two exposed global latencies per 64-row tile).  Here a warp owns 128 features
// (one float4 per lane) and walks rows r = warp, warp + 8, ...: 8 independent 16-byte loads in flight per
// lane, dY rows read as warp-uniform __ldg (L1-resident, 36 B per row), no shared-memory staging and no
// barrier in the row loop.  The 4 warps of a block are combined through shared memory, one atomic per
// (k, n) per block; the bias gradient (column sums of dY) falls out of the same pass.
template <int NMAX>
__global__ void __launch_bounds__(128)
wgrad_skinny_v4_kernel(long M, int K, int N, long rpb, const float *__restrict__ A, int lda,
                       const float *__restrict__ a_scale, const float *__restrict__ a_shift, int a_relu,
                       const float *__restrict__ dY, float *__restrict__ dW, float *__restrict__ db) {
    pdl_enter();
    __shared__ float red[4][NMAX][132];  // [warp][n][k within the 128-feature block] (+4: bank spread)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.y * 128 + lane * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_scale) {
        float *s4 = reinterpret_cast<float *>(&sc), *h4 = reinterpret_cast<float *>(&sh);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + j < K) {
                s4[j] = __ldg(a_scale + k0 + j);
                h4[j] = __ldg(a_shift + k0 + j);
            }
    }
    float acc[NMAX][4];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    float bsum = 0.f;  // lane n < N: column sum of dY over this warp's rows
    const long r0 = (long)blockIdx.x * rpb;
    const long r1 = r0 + rpb < M ? r0 + rpb : M;
    const bool full4 = k0 + 3 < K;
#pragma unroll 8
    for (long r = r0 + warp; r < r1; r += 4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *ap = A + r * lda + k0;
        if (full4) a = __ldg(reinterpret_cast<const float4 *>(ap));
        else {
            if (k0 < K) a.x = __ldg(ap);
            if (k0 + 1 < K) a.y = __ldg(ap + 1);
            if (k0 + 2 < K) a.z = __ldg(ap + 2);
        }
        if (a_scale) {
            a.x = __fmaf_rn(a.x, sc.x, sh.x); a.y = __fmaf_rn(a.y, sc.y, sh.y);
            a.z = __fmaf_rn(a.z, sc.z, sh.z); a.w = __fmaf_rn(a.w, sc.w, sh.w);
            if (a_relu) {
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
            }
            if (k0 >= K) a.x = 0.f;
            if (k0 + 1 >= K) a.y = 0.f;
            if (k0 + 2 >= K) a.z = 0.f;
            if (k0 + 3 >= K) a.w = 0.f;
        }
        const float *gy = dY + r * N;
        const float mine = lane < N ? __ldg(gy + lane) : 0.f;
        bsum += mine;
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
            const float g = __shfl_sync(0xFFFFFFFFu, mine, n);
            acc[n][0] = __fmaf_rn(a.x, g, acc[n][0]);
            acc[n][1] = __fmaf_rn(a.y, g, acc[n][1]);
            acc[n][2] = __fmaf_rn(a.z, g, acc[n][2]);
            acc[n][3] = __fmaf_rn(a.w, g, acc[n][3]);
        }
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
        *reinterpret_cast<float4 *>(&red[warp][n][lane * 4]) = make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]);
    __syncthreads();
    // the block's 128 x N slice of dW is one contiguous run of floats: four consecutive ones per thread go out
    // as ONE vector reduction (the same ~1k addresses take the sums of every block: the L2 atomic units, not
    // the row loop, bounded this kernel)
    float *base = dW + (long)blockIdx.y * 128 * N;
    const int kvalid = K - blockIdx.y * 128 < 128 ? K - blockIdx.y * 128 : 128;
    const int fvalid = kvalid * N;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    for (int q = threadIdx.x; q < N * 32; q += 128) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = 4 * q + j, kk = f / N, n = f - kk * N;
            v[j] = (red[0][n][kk] + red[1][n][kk]) + (red[2][n][kk] + red[3][n][kk]);
        }
        if (vec_ok && 4 * q + 3 < fvalid) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * q), "f"(v[0]), "f"(v[1]),
                         "f"(v[2]), "f"(v[3])
                         : "memory");
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * q + j < fvalid) atomicAdd(base + 4 * q + j, v[j]);
        }
    }
    if (db && blockIdx.y == 0) {
        // one atomic per column and block
        __syncthreads();
        if (lane < N) red[warp][0][lane] = bsum;
        __syncthreads();
        if (threadIdx.x < N)
            atomicAdd(db + threadIdx.x, (red[0][0][threadIdx.x] + red[1][0][threadIdx.x]) +
                                            (red[2][0][threadIdx.x] + red[3][0][threadIdx.x]));
    }
}

// dW[K,N] += f(A)^T dY for narrow outputs (N <= 16: the 9-class head).  Thread = feature k with N
// accumulators in registers; the dY rows of a 64-row tile are broadcast from shared memory, the A
// column reads are coalesced across the block.  One fp32 atomic per (k, n) per block.
template <int NMAX>
__global__ void __launch_bounds__(128)
wgrad_skinny_kernel(long M, int K, int N, long rpb, const float *__restrict__ A, int lda,
                    const float *__restrict__ a_scale, const float *__restrict__ a_shift, int a_relu,
                    const float *__restrict__ dY, float *__restrict__ dW) {
    pdl_enter();
    __shared__ float sdy[64 * NMAX];
    const int k = blockIdx.y * 128 + threadIdx.x;
    const bool kok = k < K;
    const float sc = (a_scale && kok) ? __ldg(a_scale + k) : 1.f;
    const float sh = (a_scale && kok) ? __ldg(a_shift + k) : 0.f;
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    for (int e = threadIdx.x; e < 64 * NMAX; e += 128) sdy[e] = 0.f;
    const long r0 = (long)blockIdx.x * rpb;
    const long r1 = r0 + rpb < M ? r0 + rpb : M;
    for (long rt = r0; rt < r1; rt += 64) {
        const int nr = (int)((r1 - rt) < 64 ? (r1 - rt) : 64);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * N; e += 128)
            sdy[(e / N) * NMAX + (e % N)] = __ldg(dY + rt * N + e);
        __syncthreads();
        const float *ap = A + rt * lda + k;
#pragma unroll 8
        for (int r = 0; r < nr; ++r) {
            float a = kok ? __ldg(ap + (long)r * lda) : 0.f;
            if (a_scale) {
                a = __fmaf_rn(a, sc, sh);
                if (a_relu) a = fmaxf(a, 0.f);
            }
#pragma unroll
            for (int n = 0; n < NMAX; ++n) acc[n] = __fmaf_rn(a, sdy[r * NMAX + n], acc[n]);
        }
    }
    if (kok) {
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) atomicAdd(dW + (long)k * N + n, acc[n]);
    }
}

__global__ void colsum_kernel(long M, int N, long rpb, const float *__restrict__ X,
                              float *__restrict__ out) {
    pdl_enter();
    const Slab s = make_slab(M, N, rpb);
    if (!s.active) return;
    for (int c = s.cx; c < N; c += s.lanes) {
        double a = 0.0;
        for (long r = s.r0 + s.ry; r < s.r1; r += s.rpp) a += (double)__ldg(X + r * N + c);
        atomicAdd(out + c, (float)a);
    }
}

static inline bool vec4_ok(int N, int ldz, const void *a, const void *b) {
    return (N % 4 == 0) && (ldz % 4 == 0) &&
           ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

static inline int grid_for(long total, int threads) {
    long blocks = ceil_div<long>(total, threads);
    long cap = 148L * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}
// rows per block for the slab kernels: aim at ~148*8 blocks, at least 32 rows each
static inline long slab_rows(long M, int *blocks) {
    long rpb = ceil_div<long>(M, 148L * 8);
    if (rpb < 32) rpb = 32;
    *blocks = (int)ceil_div<long>(M, rpb);
    return rpb;
}

}  // namespace pn2

using namespace pn2;

static bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PN2_DISABLE_TC");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

PN2_API int pn2_linear_fwd(long M, int K, int N, const float *A, int lda, const float *a_scale,
                           const float *a_shift, int a_relu, const float *W, const float *bias,
                           float *Y, double *stats, void *ws, long ws_bytes, int mode,
                           pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && K > 0 && N > 0 && lda >= K && M < (1L << 31));
    const bool image_ready = mode >= PN2_GEMM_IMAGE_READY - 1;  // mode + PN2_GEMM_IMAGE_READY
    if (image_ready) mode -= PN2_GEMM_IMAGE_READY;
    PN2_REQUIRE(mode >= -1 && mode <= 1);
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(A);
    PN2_REQUIRE_PTR(W);
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE((a_scale == nullptr) == (a_shift == nullptr));
    cudaStream_t st = as_stream(s);
    if (mode == 1 || (mode == -1 && tc_enabled())) {
        int rc = tc_linear_fwd(M, K, N, A, lda, a_scale, a_shift, a_relu, W, bias, Y, stats,
                               static_cast<float *>(ws), ws_bytes > 0 ? (size_t)ws_bytes : 0, image_ready, nullptr,
                               st);
        if (rc != PN2_EUNSUPPORTED || mode == 1) return rc;
    }
    return launch_gemm<true, true, false>((int)M, N, K, A, lda, 1, W, N, 1, a_scale, a_shift,
                                          a_relu, bias, Y, N, stats, 1, st);
}

PN2_API int pn2_linear_fwd_bn(long M, int K, int N, const float *A, int lda, const float *a_scale,
                              const float *a_shift, int a_relu, const float *W, const float *bias, float *Y,
                              double *stats, const pn2_bn_finalize *fin, void *ws, long ws_bytes, int mode,
                              pn2_stream_t s) {
    PN2_REQUIRE(M > 0 && K > 0 && N > 0 && lda >= K && M < (1L << 31));
    const bool image_ready = mode >= PN2_GEMM_IMAGE_READY - 1;
    if (image_ready) mode -= PN2_GEMM_IMAGE_READY;
    PN2_REQUIRE(mode >= -1 && mode <= 1);
    PN2_REQUIRE_PTR(A);
    PN2_REQUIRE_PTR(W);
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(stats);
    PN2_REQUIRE_PTR(fin);
    PN2_REQUIRE_PTR(fin->gamma);
    PN2_REQUIRE_PTR(fin->beta);
    PN2_REQUIRE_PTR(fin->scale);
    PN2_REQUIRE_PTR(fin->shift);
    PN2_REQUIRE_PTR(fin->saved);
    PN2_REQUIRE_PTR(fin->counter);
    PN2_REQUIRE((fin->moving_mean == nullptr) == (fin->moving_var == nullptr));
    PN2_REQUIRE((a_scale == nullptr) == (a_shift == nullptr));
    cudaStream_t st = as_stream(s);
    if (mode == 1 || (mode == -1 && tc_enabled())) {
        int rc = tc_linear_fwd(M, K, N, A, lda, a_scale, a_shift, a_relu, W, bias, Y, stats,
                               static_cast<float *>(ws), ws_bytes > 0 ? (size_t)ws_bytes : 0, image_ready, fin, st);
        if (rc != PN2_EUNSUPPORTED || mode == 1) return rc;
    }
    int rc = launch_gemm<true, true, false>((int)M, N, K, A, lda, 1, W, N, 1, a_scale, a_shift, a_relu, bias, Y, N,
                                            stats, 1, st);
    if (rc) return rc;
    return pn2_bn_train_finalize(N, M, stats, fin->gamma, fin->beta, fin->eps, fin->decay, fin->unbiased_moving,
                                 fin->moving_mean, fin->moving_var, fin->scale, fin->shift, fin->saved, s);
}

PN2_API long pn2_linear_image_bytes(int K, int N, int dgrad) {
    if (K <= 0 || N <= 0) return 0;
    return (long)tc_image_bytes_for(K, N, dgrad != 0);
}

PN2_API int pn2_linear_image_describe(int K, int N, int dgrad, const float *W, float *image,
                                      pn2_linear_image *out) {
    PN2_REQUIRE(K > 0 && N > 0);
    PN2_REQUIRE_PTR(W);
    PN2_REQUIRE_PTR(image);
    PN2_REQUIRE_PTR(out);
    PN2_REQUIRE((reinterpret_cast<uintptr_t>(image) & 127) == 0);
    return tc_describe_image(K, N, dgrad != 0, W, image, out);
}

PN2_API int pn2_linear_prepare(int count, const pn2_linear_image *table_dev, pn2_stream_t s) {
    PN2_REQUIRE(count >= 0 && count <= 65535);
    if (count == 0) return PN2_OK;
    PN2_REQUIRE_PTR(table_dev);
    return tc_prepare_images(count, table_dev, as_stream(s));
}

PN2_API long pn2_linear_workspace_bytes(int K, int N) {
    if (K <= 0 || N <= 0) return 0;
    return (long)tc_workspace_bytes(K, N);
}

PN2_API int pn2_linear_dgrad(long M, int K, int N, const float *dY, const float *W, float *dX,
                             int ldx, void *ws, long ws_bytes, int mode, pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && K > 0 && N > 0 && ldx >= K && M < (1L << 31));
    const bool image_ready = mode >= PN2_GEMM_IMAGE_READY - 1;  // mode + PN2_GEMM_IMAGE_READY
    if (image_ready) mode -= PN2_GEMM_IMAGE_READY;
    PN2_REQUIRE(mode >= -1 && mode <= 1);
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dY);
    PN2_REQUIRE_PTR(W);
    PN2_REQUIRE_PTR(dX);
    cudaStream_t st = as_stream(s);
    if (mode == 1 || (mode == -1 && tc_enabled())) {
        int rc = tc_linear_dgrad(M, K, N, dY, W, dX, ldx, static_cast<float *>(ws),
                                 ws_bytes > 0 ? (size_t)ws_bytes : 0, image_ready, st);
        if (rc != PN2_EUNSUPPORTED || mode == 1) return rc;
    }
    // C[M,K] = sum_n dY(m,n) * W(k,n):  M'=M, N'=K, K'=N ; B(k'=n, n'=k) = W + k*N + n
    return launch_gemm<true, false, false>((int)M, K, N, dY, N, 1, W, 1, N, nullptr, nullptr, 0,
                                           nullptr, dX, ldx, nullptr, 1, st);
}

PN2_API int pn2_linear_wgrad(long M, int K, int N, const float *A, int lda, const float *a_scale,
                             const float *a_shift, int a_relu, const float *dY, float *dW,
                             float *db, int mode, pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && K > 0 && N > 0 && lda >= K && M < (1L << 31));
    PN2_REQUIRE(mode >= -1 && mode <= 1);
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(A);
    PN2_REQUIRE_PTR(dY);
    PN2_REQUIRE_PTR(dW);
    PN2_REQUIRE((a_scale == nullptr) == (a_shift == nullptr));
    cudaStream_t st = as_stream(s);
    int rc = PN2_EUNSUPPORTED;
    // (the register-tiled streaming kernel of pn2_wgrad_rt.cuh measured slower than the split-K tile
    //  kernel on every model shape -- 174 vs 101 us at 524288x32x32 -- and is kept for PN2_WGRAD_RT=1)
    static const bool use_rt = getenv("PN2_WGRAD_RT") && getenv("PN2_WGRAD_RT")[0] == '1';
    if (use_rt && mode != 1) rc = wgrad_rt(M, K, N, A, lda, a_scale, a_shift, a_relu, dY, dW, st);
    int k_done = 0;
    if (rc == PN2_EUNSUPPORTED && mode != 1 && N <= 16) {  // narrow output: feature-per-thread kernel
        static const bool old_skinny = getenv("PN2_WGRAD_SKINNY_V1") && getenv("PN2_WGRAD_SKINNY_V1")[0] == '1';
        const bool v4 = !old_skinny && (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
        if (v4) {  // warp per 128 features, float4 rows, bias gradient in the same pass
            static const int bps = getenv("PN2_SKINNY_BPS") ? atoi(getenv("PN2_SKINNY_BPS")) : 3;
            long rpb = ceil_div<long>(M, 148L * (bps > 0 ? bps : 3));
            rpb = ceil_div<long>(rpb, 8L) * 8;
            dim3 grid((unsigned)ceil_div<long>(M, rpb), (unsigned)ceil_div(K, 128));
            launch_k(wgrad_skinny_v4_kernel<16>, grid, 128, 0, st, M, K, N, rpb, A, lda, a_scale, a_shift, a_relu, dY,
                                                             dW, db);
            rc = finish_launch();
            if (rc) return rc;
            db = nullptr;  // done
        } else {
            // ~7 blocks of 128 threads per SM, 8 independent row loads in flight per thread
            long rpb = ceil_div<long>(M, 148L * 7);
            rpb = ceil_div<long>(rpb, 64L) * 64;
            dim3 grid((unsigned)ceil_div<long>(M, rpb), (unsigned)ceil_div(K, 128));
            launch_k(wgrad_skinny_kernel<16>, grid, 128, 0, st, M, K, N, rpb, A, lda, a_scale, a_shift, a_relu, dY, dW);
            rc = finish_launch();
            if (rc) return rc;
        }
        k_done = K;
    }
    if (rc == PN2_EUNSUPPORTED && (mode == 1 || (mode == -1 && tc_enabled())))
        rc = tc_linear_wgrad(M, K, N, A, lda, a_scale, a_shift, a_relu, dY, dW, mode == 1, &k_done, st);
    if (rc == PN2_OK && k_done > 0 && k_done < K) {
        // feature tail [k_done, K) on the fp32 split-K kernel
        const int kt = K - k_done;
        int sp = (int)ceil_div<long>(148L * 4, (long)ceil_div(N, N > 64 ? 128 : (N > 32 ? 64 : 32)));
        const long ms = ceil_div<long>(M, 256);
        if (sp > ms) sp = (int)ms;
        if (sp < 1) sp = 1;
        rc = launch_gemm<false, true, true>(kt, N, M, A + k_done, 1, lda, dY, N, 1,
                                            a_scale ? a_scale + k_done : nullptr,
                                            a_shift ? a_shift + k_done : nullptr, a_relu, nullptr,
                                            dW + (long)k_done * N, N, nullptr, sp, st);
    }
    if (rc != PN2_EUNSUPPORTED || mode == 1) {
        if (rc == PN2_OK && db) {
            long rpb = ceil_div<long>(M, 64L);
            if (rpb < 32) rpb = 32;
            launch_k(colsum_kernel, (int)ceil_div<long>(M, rpb), 256, 0, st, M, N, rpb, dY, db);
            rc = finish_launch();
        }
        return rc;
    }
    // C[K,N] += sum_m f(A)(m,k) dY(m,n): M'=K, N'=N, K'=M ; A(m'=k, k'=m) = A + m*lda + k
    const int tiles = ceil_div(K, K > 64 ? 128 : (K > 32 ? 64 : 32)) *
                      ceil_div(N, N > 64 ? 128 : (N > 32 ? 64 : 32));
    // enough CTAs to fill the machine (measured: limiting the split count to save atomics made
    // short problems 3x slower -- parallelism over the contraction rows matters more)
    int splits = (int)ceil_div<long>(148L * 4, tiles);
    const long max_splits = ceil_div<long>(M, 256);
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    rc = launch_gemm<false, true, true>(K, N, M, A, 1, lda, dY, N, 1, a_scale, a_shift, a_relu,
                                        nullptr, dW, N, nullptr, splits, st);
    if (rc) return rc;
    if (db) {
        long rpb = ceil_div<long>(M, 64L);
        if (rpb < 32) rpb = 32;
        int blocks = (int)ceil_div<long>(M, rpb);
        launch_k(colsum_kernel, blocks, 256, 0, st, M, N, rpb, dY, db);
        rc = finish_launch();
    }
    return rc;
}

PN2_API int pn2_bn_train_finalize(int N, long M, const double *stats, const float *gamma,
                                  const float *beta, float eps, float decay, int unbiased_moving,
                                  float *moving_mean, float *moving_var, float *scale,
                                  float *shift, float *saved, pn2_stream_t s) {
    PN2_REQUIRE(N > 0 && M > 0);
    PN2_REQUIRE_PTR(stats);
    PN2_REQUIRE_PTR(gamma);
    PN2_REQUIRE_PTR(beta);
    PN2_REQUIRE_PTR(scale);
    PN2_REQUIRE_PTR(shift);
    PN2_REQUIRE_PTR(saved);
    PN2_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr));
    launch_k(bn_train_finalize_kernel, ceil_div(N, 128), 128, 0, as_stream(s), 
        N, M, stats, gamma, beta, eps, decay, unbiased_moving, moving_mean, moving_var, scale, shift,
        saved);
    return finish_launch();
}

PN2_API int pn2_bn_eval_affine(int N, const float *gamma, const float *beta,
                               const float *moving_mean, const float *moving_var, float eps,
                               float *scale, float *shift, pn2_stream_t s) {
    PN2_REQUIRE(N > 0);
    PN2_REQUIRE_PTR(gamma);
    PN2_REQUIRE_PTR(beta);
    PN2_REQUIRE_PTR(moving_mean);
    PN2_REQUIRE_PTR(moving_var);
    PN2_REQUIRE_PTR(scale);
    PN2_REQUIRE_PTR(shift);
    launch_k(bn_eval_affine_kernel, ceil_div(N, 128), 128, 0, as_stream(s), 
        N, gamma, beta, moving_mean, moving_var, eps, scale, shift);
    return finish_launch();
}

PN2_API int pn2_affine_act(long M, int N, const float *Y, const float *scale, const float *shift,
                           int relu, float *Z, int ldz, pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && N > 0 && ldz >= N);
    PN2_REQUIRE((scale == nullptr) == (shift == nullptr));
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(Z);
    const long total = M * N;
    const bool v4 = vec4_ok(N, ldz, Y, Z) &&
                    (scale == nullptr ||
                     ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0);
    if (v4)
        launch_k(affine_act_v4_kernel, grid_for(total / 4, 256), 256, 0, as_stream(s), total / 4, N / 4, Y, scale,
                                                                               shift, relu, Z, ldz);
    else
        launch_k(affine_act_kernel, grid_for(total, 256), 256, 0, as_stream(s), total, N, Y, scale, shift,
                                                                          relu, Z, ldz);
    return finish_launch();
}

PN2_API int pn2_affine_act_maxpool(long G, int ns, int N, const float *Y, const float *scale,
                                   const float *shift, int relu, float *out, int *arg,
                                   pn2_stream_t s) {
    PN2_REQUIRE(G >= 0 && ns > 0 && N > 0);
    PN2_REQUIRE((scale == nullptr) == (shift == nullptr));
    if (G == 0) return PN2_OK;
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(out);
    const long total = G * N;
    const bool v4 = vec4_ok(N, N, Y, out) && (arg == nullptr || (reinterpret_cast<uintptr_t>(arg) & 15) == 0) &&
                    (scale == nullptr ||
                     ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0);
    if (v4)
        launch_k(affine_act_maxpool_v4_kernel, grid_for(total / 4, 128), 128, 0, as_stream(s), 
            total / 4, ns, N, Y, scale, shift, relu, out, arg);
    else
        launch_k(affine_act_maxpool_kernel, grid_for(total, 128), 128, 0, as_stream(s), 
            total, ns, N, Y, scale, shift, relu, out, arg);
    return finish_launch();
}

PN2_API int pn2_bn_bwd_reduce(long M, int N, const float *dZ, int ldz, const float *Y,
                              const float *scale, const float *shift, const float *saved,
                              int relu, double *red, pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && N > 0 && ldz >= N);
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dZ);
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(scale);
    PN2_REQUIRE_PTR(shift);
    PN2_REQUIRE_PTR(saved);
    PN2_REQUIRE_PTR(red);
    int blocks;
    long rpb = slab_rows(M, &blocks);
    if (vec4_ok(N, ldz, dZ, Y)) {
        // a few resident blocks of work per SM: every block ends with 2N global fp64 atomics
        // (PN2_BNRED_BPS: blocks per SM, A/B switch; 80 registers x 256 threads allow 3.  Alone on the device 2 and 3
        //  measure the same; next to the weight-gradient stream, whose CTAs hold most of 64 SMs, 3 is 0.5 % of the
        //  step faster: 3.257 vs 3.240 ms, three runs each on one box)
        static const int bps = getenv("PN2_BNRED_BPS") ? atoi(getenv("PN2_BNRED_BPS")) : 3;
        long rpb4 = ceil_div<long>(M, 148L * (bps > 0 ? bps : 3));
        if (rpb4 < 32) rpb4 = 32;
        const int blocks4 = (int)ceil_div<long>(M, rpb4);
        launch_k(bn_bwd_reduce_v4_kernel, blocks4, 256, 2 * N * sizeof(double), as_stream(s), 
            M, N, rpb4, dZ, ldz, Y, scale, shift, saved, relu, red);
    }
    else
        launch_k(bn_bwd_reduce_kernel, blocks, 256, 2 * N * sizeof(double), as_stream(s), 
            M, N, rpb, dZ, ldz, Y, scale, shift, saved, relu, red);
    return finish_launch();
}

PN2_API int pn2_bn_bwd_apply(long M, int N, const float *dZ, int ldz, const float *Y,
                             const float *scale, const float *shift, const float *saved,
                             const float *gamma, int relu, int bn, const double *red, float *dY,
                             float *dgamma, float *dbeta, pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && N > 0 && ldz >= N);
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dZ);
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(dY);
    if (bn) {
        PN2_REQUIRE_PTR(scale);
        PN2_REQUIRE_PTR(shift);
        PN2_REQUIRE_PTR(saved);
        PN2_REQUIRE_PTR(gamma);
        PN2_REQUIRE_PTR(red);
    }
    PN2_REQUIRE((scale == nullptr) == (shift == nullptr));
    int blocks;
    long rpb = slab_rows(M, &blocks);
    if (vec4_ok(N, ldz, dZ, Y) && (reinterpret_cast<uintptr_t>(dY) & 15) == 0 &&
        (scale == nullptr || ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0))
        launch_k(bn_bwd_apply_v4_kernel, blocks, 256, 0, as_stream(s), M, N, rpb, dZ, ldz, Y, scale, shift,
                                                                 saved, gamma, relu, bn, red, dY, dgamma,
                                                                 dbeta);
    else
        launch_k(bn_bwd_apply_kernel, blocks, 256, 0, as_stream(s), M, N, rpb, dZ, ldz, Y, scale, shift, saved,
                                                              gamma, relu, bn, red, dY, dgamma, dbeta);
    return finish_launch();
}

PN2_API int pn2_bn_bwd_reduce_pool(long G, int ns, int N, const float *dOut, const int *arg,
                                   const float *Y, const float *scale, const float *shift,
                                   const float *saved, int relu, double *red, pn2_stream_t s) {
    PN2_REQUIRE(G >= 0 && ns > 0 && N > 0);
    if (G == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dOut);
    PN2_REQUIRE_PTR(arg);
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(scale);
    PN2_REQUIRE_PTR(shift);
    PN2_REQUIRE_PTR(saved);
    PN2_REQUIRE_PTR(red);
    // ~2048 (group, column) elements per block, at most four blocks per SM
    long blocks = ceil_div<long>(G * (long)N, 2048L);
    if (blocks > 148L * 4) blocks = 148L * 4;
    if (blocks > G) blocks = G;
    if (blocks < 1) blocks = 1;
    launch_k(bn_bwd_reduce_pool_kernel, (int)blocks, 256, 2 * N * sizeof(double), as_stream(s), 
        G, ns, N, dOut, arg, Y, scale, shift, saved, relu, red);
    return finish_launch();
}

PN2_API int pn2_bn_bwd_apply_pool(long G, int ns, int N, const float *dOut, const int *arg,
                                  const float *Y, const float *scale, const float *shift,
                                  const float *saved, const float *gamma, int relu, int bn,
                                  const double *red, float *dY, float *dgamma, float *dbeta,
                                  pn2_stream_t s) {
    PN2_REQUIRE(G >= 0 && ns > 0 && N > 0);
    if (G == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dOut);
    PN2_REQUIRE_PTR(arg);
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(dY);
    if (bn) {
        PN2_REQUIRE_PTR(scale);
        PN2_REQUIRE_PTR(shift);
        PN2_REQUIRE_PTR(saved);
        PN2_REQUIRE_PTR(gamma);
        PN2_REQUIRE_PTR(red);
    }
    PN2_REQUIRE((scale == nullptr) == (shift == nullptr));
    const bool v4 = (N % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(dOut) | reinterpret_cast<uintptr_t>(arg) |
                      reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(dY)) & 15) == 0;
    if (v4) {
        long gpb = ceil_div<long>(G, 148L * 8);
        if (gpb < 1) gpb = 1;
        launch_k(bn_bwd_apply_pool_v4_kernel, (int)ceil_div<long>(G, gpb), 256, 0, as_stream(s), 
            G, ns, N, gpb, dOut, arg, Y, scale, shift, saved, gamma, relu, bn, red, dY, dgamma, dbeta);
        return finish_launch();
    }
    int blocks;
    long rpb = slab_rows(G * ns, &blocks);
    launch_k(bn_bwd_apply_pool_kernel, blocks, 256, 0, as_stream(s), G, ns, N, rpb, dOut, arg, Y, scale,
                                                               shift, saved, gamma, relu, bn, red, dY,
                                                               dgamma, dbeta);
    return finish_launch();
}

PN2_API int pn2_pool_weights(long G, int ns, const float *grouped_xyz, int ld, float *w, pn2_stream_t s) {
    PN2_REQUIRE(G >= 0 && ns > 0 && ld >= 3);
    if (G == 0) return PN2_OK;
    PN2_REQUIRE_PTR(grouped_xyz);
    PN2_REQUIRE_PTR(w);
    launch_k(pool_weights_kernel, grid_for(G, 128), 128, 0, as_stream(s), G, ns, grouped_xyz, ld, w);
    return finish_launch();
}

PN2_API int pn2_group_pool(long G, int ns, int N, const float *X, const float *w, int mode, float *out,
                           int *arg, pn2_stream_t s) {
    PN2_REQUIRE(G >= 0 && ns > 0 && N > 0 && mode >= 1 && mode <= 3);
    if (G == 0) return PN2_OK;
    PN2_REQUIRE_PTR(X);
    PN2_REQUIRE_PTR(out);
    if (mode == 2) PN2_REQUIRE_PTR(w);
    if (mode == 3) PN2_REQUIRE_PTR(arg);
    launch_k(group_pool_kernel, grid_for(G * N, 128), 128, 0, as_stream(s), G * N, ns, N, X, w, mode, out, arg);
    return finish_launch();
}

PN2_API int pn2_group_pool_grad(long G, int ns, int N, const float *dOut, const float *w, const int *arg,
                                int mode, float *dX, pn2_stream_t s) {
    PN2_REQUIRE(G >= 0 && ns > 0 && N > 0 && mode >= 1 && mode <= 3);
    if (G == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dOut);
    PN2_REQUIRE_PTR(dX);
    if (mode == 2) PN2_REQUIRE_PTR(w);
    if (mode == 3) PN2_REQUIRE_PTR(arg);
    launch_k(group_pool_grad_kernel, grid_for(G * ns * N, 256), 256, 0, as_stream(s), G * ns * N, ns, N, dOut, w, arg,
                                                                                mode, dX);
    return finish_launch();
}

PN2_API int pn2_relu_mask(long M, int N, const float *Y, const float *scale, const float *shift,
                          unsigned char *mask, pn2_stream_t s) {
    PN2_REQUIRE(M >= 0 && N > 0);
    PN2_REQUIRE((scale == nullptr) == (shift == nullptr));
    if (M == 0) return PN2_OK;
    PN2_REQUIRE_PTR(Y);
    PN2_REQUIRE_PTR(mask);
    launch_k(relu_mask_kernel, grid_for(M * N, 256), 256, 0, as_stream(s), M * N, N, Y, scale, shift, mask);
    return finish_launch();
}

PN2_API int pn2_dropout(long n, const float *x, float keep_prob, unsigned long long seed,
                        const unsigned long long *seed_dev, float *out, pn2_stream_t s) {
    PN2_REQUIRE(n >= 0 && keep_prob > 0.f && keep_prob <= 1.f);
    if (n == 0) return PN2_OK;
    PN2_REQUIRE_PTR(x);
    PN2_REQUIRE_PTR(out);
    if ((n % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
        launch_k(dropout_v4_kernel, grid_for(n / 4, 256), 256, 0, as_stream(s), 
            n / 4, reinterpret_cast<const float4 *>(x), keep_prob, 1.0f / keep_prob, seed, seed_dev,
            reinterpret_cast<float4 *>(out));
    else
        launch_k(dropout_kernel, grid_for(n, 256), 256, 0, as_stream(s), n, x, keep_prob, 1.0f / keep_prob,
                                                                   seed, seed_dev, out);
    return finish_launch();
}

PN2_API int pn2_dropout_mask(long n, float keep_prob, unsigned long long seed,
                             unsigned char *mask, pn2_stream_t s) {
    PN2_REQUIRE(n >= 0 && keep_prob > 0.f && keep_prob <= 1.f);
    if (n == 0) return PN2_OK;
    PN2_REQUIRE_PTR(mask);
    launch_k(dropout_mask_kernel, grid_for(n, 256), 256, 0, as_stream(s), n, keep_prob, seed, mask);
    return finish_launch();
}

PN2_API int pn2_softmax_ce_reduce(long rows, int C, const float *logits, const int *labels,
                                  const float *weights, double *acc, pn2_stream_t s) {
    PN2_REQUIRE(rows >= 0 && C > 0);
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE_PTR(logits);
    PN2_REQUIRE_PTR(labels);
    PN2_REQUIRE_PTR(acc);
    launch_k(softmax_ce_reduce_kernel, grid_for(rows, 256), 256, 0, as_stream(s), rows, C, logits, labels,
                                                                           weights, acc);
    return finish_launch();
}

PN2_API int pn2_softmax_ce_grad(long rows, int C, const float *logits, const int *labels,
                                const float *weights, const double *acc, float gscale,
                                const float *gscale_dev, float *loss, float *dlogits,
                                pn2_stream_t s) {
    PN2_REQUIRE(rows >= 0 && C > 0);
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE_PTR(logits);
    PN2_REQUIRE_PTR(labels);
    PN2_REQUIRE_PTR(acc);
    launch_k(softmax_ce_grad_kernel, grid_for(rows, 256), 256, 0, as_stream(s), 
        rows, C, logits, labels, weights, acc, gscale, gscale_dev, loss, dlogits);
    return finish_launch();
}

PN2_API int pn2_adam_step(long n, float *p, const float *g, float *m, float *v, float lr,
                          float beta1, float beta2, float eps, int t, float gscale,
                          pn2_stream_t s) {
    PN2_REQUIRE(n >= 0 && t >= 1);
    if (n == 0) return PN2_OK;
    PN2_REQUIRE_PTR(p);
    PN2_REQUIRE_PTR(g);
    PN2_REQUIRE_PTR(m);
    PN2_REQUIRE_PTR(v);
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t));
    launch_k(adam_kernel, grid_for(n, 256), 256, 0, as_stream(s), n, p, g, m, v, (float)lr_t, beta1, beta2,
                                                            eps, gscale);
    return finish_launch();
}
