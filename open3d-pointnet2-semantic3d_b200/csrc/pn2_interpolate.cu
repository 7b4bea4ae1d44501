// pn2_interpolate.cu -- three_nn, inverse-distance weights, three_interpolate(+grad), sm_100a.
//
// Replaces tf_ops/tf_interpolate.cpp:213-243, 307-330, 397-421 of the reference, which are
// single-threaded CPU loops (three_nn through an Open3D KD-tree) that force device<->host
// copies in the middle of the network.  Here:
//   three_nn: brute force on the GPU, one thread per query, the known points staged once per
//     CTA into shared memory already promoted to fp64 (the reference searches on fp64 copies of
//     the fp32 coordinates, tf_interpolate.cpp:20-28); d = (dx*dx + dy*dy) + dz*dz with
//     explicit round-to-nearest mul/add (no FMA contraction), top-3 kept ascending with strict
//     '<' so equal distances resolve to the lowest index; results cast to fp32 like the
//     reference's double->float assignment (:238-240).
//   three_interpolate: (p1*w1 + p2*w2) + p3*w3 evaluated with separate mul/add (the reference
//     CPU build has no FMA contraction), vectorised over channels.
#include <stdlib.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;  // known points per stage: 24 KB of doubles

__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                float *__restrict__ dist, int *__restrict__ idx) {
    pdl_enter();
    __shared__ __align__(16) double kx[NN_TILE], ky[NN_TILE], kz[NN_TILE];
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool valid = j < n;
    double qx = 0, qy = 0, qz = 0;
    if (valid) {
        const float *q = xyz1 + ((size_t)cloud * n + j) * 3;
        qx = (double)__ldg(q);
        qy = (double)__ldg(q + 1);
        qz = (double)__ldg(q + 2);
    }
    const double INF = __longlong_as_double(0x7FF0000000000000LL);
    double b1 = INF, b2 = INF, b3 = INF;
    int i1 = 0, i2 = 0, i3 = 0;
    const float *known = xyz2 + (size_t)cloud * m * 3;
    for (int base = 0; base < m; base += NN_TILE) {
        const int cnt = min(NN_TILE, m - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += NN_THREADS) {
            int k = e / 3, c = e - k * 3;
            double v = (double)__ldg(known + (size_t)base * 3 + e);
            (c == 0 ? kx : (c == 1 ? ky : kz))[k] = v;
        }
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                double dx = __dsub_rn(qx, kx[k]);
                double dy = __dsub_rn(qy, ky[k]);
                double dz = __dsub_rn(qz, kz[k]);
                double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)),
                                     __dmul_rn(dz, dz));
                if (d < b3) {
                    const int kk = base + k;
                    if (d < b1) {
                        b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk;
                    } else if (d < b2) {
                        b3 = b2; i3 = i2; b2 = d; i2 = kk;
                    } else {
                        b3 = d; i3 = kk;
                    }
                }
            }
        }
    }
    if (valid) {
        size_t o = ((size_t)cloud * n + j) * 3;
        dist[o] = (float)b1;
        dist[o + 1] = (float)b2;
        dist[o + 2] = (float)b3;
        idx[o] = i1;
        idx[o + 1] = i2;
        idx[o + 2] = i3;
    }
}

// pointnet_util.py:300-303 (fp32, true division):
//   d = max(d,1e-10); norm = (1/d0 + 1/d1) + 1/d2; w = (1/d)/norm
__global__ void fp_weights_kernel(long rows, const float *__restrict__ dist,
                                  float *__restrict__ weight) {
    pdl_enter();
    for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < rows;
         r += (long)gridDim.x * blockDim.x) {
        float d0 = fmaxf(__ldg(dist + r * 3), 1e-10f);
        float d1 = fmaxf(__ldg(dist + r * 3 + 1), 1e-10f);
        float d2 = fmaxf(__ldg(dist + r * 3 + 2), 1e-10f);
        float r0 = __fdiv_rn(1.0f, d0), r1 = __fdiv_rn(1.0f, d1), r2 = __fdiv_rn(1.0f, d2);
        float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        weight[r * 3] = __fdiv_rn(r0, norm);
        weight[r * 3 + 1] = __fdiv_rn(r1, norm);
        weight[r * 3 + 2] = __fdiv_rn(r2, norm);
    }
}

__device__ __forceinline__ float blend3(float p1, float p2, float p3, float w1, float w2,
                                        float w3) {
    return __fadd_rn(__fadd_rn(__fmul_rn(p1, w1), __fmul_rn(p2, w2)), __fmul_rn(p3, w3));
}

template <bool VEC4>
__global__ void three_interpolate_kernel(int m, int c, int n, long total,
                                         const float *__restrict__ points,
                                         const int *__restrict__ idx,
                                         const float *__restrict__ weight,
                                         float *__restrict__ out, int ldo) {
    pdl_enter();
    const int cv = VEC4 ? c / 4 : c;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / cv;  // cloud*n + j
        int l = (int)(e - row * cv);
        long cloud = row / n;
        const int i1 = __ldg(idx + row * 3), i2 = __ldg(idx + row * 3 + 1),
                  i3 = __ldg(idx + row * 3 + 2);
        const float w1 = __ldg(weight + row * 3), w2 = __ldg(weight + row * 3 + 1),
                    w3 = __ldg(weight + row * 3 + 2);
        const float *p = points + cloud * m * c;
        if (VEC4) {
            float4 a = __ldg(reinterpret_cast<const float4 *>(p + (size_t)i1 * c) + l);
            float4 b = __ldg(reinterpret_cast<const float4 *>(p + (size_t)i2 * c) + l);
            float4 d = __ldg(reinterpret_cast<const float4 *>(p + (size_t)i3 * c) + l);
            float4 o;
            o.x = blend3(a.x, b.x, d.x, w1, w2, w3);
            o.y = blend3(a.y, b.y, d.y, w1, w2, w3);
            o.z = blend3(a.z, b.z, d.z, w1, w2, w3);
            o.w = blend3(a.w, b.w, d.w, w1, w2, w3);
            *reinterpret_cast<float4 *>(out + row * ldo + l * 4) = o;
        } else {
            out[row * ldo + l] = blend3(__ldg(p + (size_t)i1 * c + l), __ldg(p + (size_t)i2 * c + l),
                                        __ldg(p + (size_t)i3 * c + l), w1, w2, w3);
        }
    }
}

__global__ void three_interpolate_grad_kernel(int m, int c, int n, long total,
                                              const float *__restrict__ grad_out, int ldg,
                                              const int *__restrict__ idx,
                                              const float *__restrict__ weight,
                                              float *__restrict__ grad_points) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / c;
        int l = (int)(e - row * c);
        long cloud = row / n;
        float g = __ldg(grad_out + row * ldg + l);
        float *gp = grad_points + cloud * m * c;
        atomicAdd(gp + (size_t)__ldg(idx + row * 3) * c + l, __fmul_rn(g, __ldg(weight + row * 3)));
        atomicAdd(gp + (size_t)__ldg(idx + row * 3 + 1) * c + l,
                  __fmul_rn(g, __ldg(weight + row * 3 + 1)));
        atomicAdd(gp + (size_t)__ldg(idx + row * 3 + 2) * c + l,
                  __fmul_rn(g, __ldg(weight + row * 3 + 2)));
    }
}

// c % 4 == 0: one thread per (row, 4 consecutive channels), three red.global.add.v4.f32 instead of twelve
// scalar atomics (the source row pitch may be odd -- 131 at FP4 -- so the four gradients load as scalars)
__global__ void three_interpolate_grad_v4_kernel(int m, int c, int n, long total4,
                                                 const float *__restrict__ grad_out, int ldg,
                                                 const int *__restrict__ idx,
                                                 const float *__restrict__ weight,
                                                 float *__restrict__ grad_points) {
    pdl_enter();
    const int c4 = c >> 2;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4;
         e += (long)gridDim.x * blockDim.x) {
        const long row = e / c4;
        const int l = 4 * (int)(e - row * c4);
        const long cloud = row / n;
        const float *g = grad_out + row * ldg + l;
        const float g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2), g3 = __ldg(g + 3);
        float *gp = grad_points + cloud * m * c + l;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const float wt = __ldg(weight + row * 3 + t);
            float *dst = gp + (size_t)__ldg(idx + row * 3 + t) * c;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__fmul_rn(g0, wt)),
                         "f"(__fmul_rn(g1, wt)), "f"(__fmul_rn(g2, wt)), "f"(__fmul_rn(g3, wt))
                         : "memory");
        }
    }
}

__global__ void copy_cols_kernel(long rows, int cols, long total, const float *__restrict__ src,
                                 int lds, float *__restrict__ dst, int ldd, int accumulate) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long r = e / cols;
        int j = (int)(e - r * cols);
        float v = __ldg(src + r * lds + j);
        if (accumulate) dst[r * ldd + j] += v;
        else dst[r * ldd + j] = v;
    }
}

__global__ void copy_cols_v4_kernel(unsigned rows, unsigned cols4, const float4 *__restrict__ src, unsigned lds4,
                                    float4 *__restrict__ dst, unsigned ldd4, int accumulate) {
    pdl_enter();
    const unsigned total = rows * cols4;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const unsigned r = e / cols4, j = e - r * cols4;
        const float4 v = __ldg(src + (size_t)r * lds4 + j);
        float4 *d = dst + (size_t)r * ldd4 + j;
        if (accumulate) {
            float4 o = *d;
            o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
            *d = o;
        } else {
            *d = v;
        }
    }
}

// ---- interpolate_label_with_color: kNN label vote (tf_interpolate.cpp:71-115) -------------
// The step right after the network at inference: every point of the full-resolution cloud takes
// the label that wins among its knn nearest predicted (sparse) points.  The reference builds an
// Open3D KD-tree on the CPU and searches it per dense point; here it is the same brute-force
// fp64 scan as three_nn (same staging, same non-contracted distance expression, strict '<' so
// equal distances keep the lowest index), with the k best candidates of a thread in a register
// array (KMAX is a compile-time bound, all indexing is unrolled) and the vote done in place:
// walking the neighbours from nearest to farthest, the label whose running count first exceeds
// the best count so far wins (:97-106) -- O(k^2) comparisons instead of a hash map.
constexpr int KV_THREADS = 256;

__constant__ unsigned char kLabelColors[9][3] = {
    {255, 255, 255}, {0, 0, 255}, {128, 0, 0}, {255, 0, 255}, {0, 128, 0},
    {255, 0, 0},     {128, 0, 128}, {0, 0, 128}, {128, 128, 0}};  // tf_interpolate.cpp:46-48

template <int KMAX>
__global__ void __launch_bounds__(KV_THREADS)
knn_vote_kernel(int ns, int nd, int k, const float *__restrict__ sparse,
                const int *__restrict__ labels, const float *__restrict__ dense,
                int *__restrict__ out_labels, unsigned char *__restrict__ out_colors) {
    pdl_enter();
    __shared__ __align__(16) double kx[NN_TILE], ky[NN_TILE], kz[NN_TILE];
    const long j = (long)blockIdx.x * KV_THREADS + threadIdx.x;
    const bool valid = j < nd;
    double qx = 0, qy = 0, qz = 0;
    if (valid) {
        const float *q = dense + (size_t)j * 3;
        qx = (double)__ldg(q);
        qy = (double)__ldg(q + 1);
        qz = (double)__ldg(q + 2);
    }
    const double INF = __longlong_as_double(0x7FF0000000000000LL);
    double bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; ++s) {
        bd[s] = INF;
        bi[s] = -1;
    }
    double worst = INF;  // bd[k-1]: a candidate must beat it to enter the list
    for (int base = 0; base < ns; base += NN_TILE) {
        const int cnt = min(NN_TILE, ns - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += KV_THREADS) {
            int p = e / 3, c = e - p * 3;
            double v = (double)__ldg(sparse + (size_t)base * 3 + e);
            (c == 0 ? kx : (c == 1 ? ky : kz))[p] = v;
        }
        __syncthreads();
        if (valid) {
#pragma unroll 2
            for (int p = 0; p < cnt; ++p) {
                double dx = __dsub_rn(qx, kx[p]);
                double dy = __dsub_rn(qy, ky[p]);
                double dz = __dsub_rn(qz, kz[p]);
                double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)),
                                     __dmul_rn(dz, dz));
                if (d < worst) {
                    // replace the current k-th, then bubble towards the front (strict '<':
                    // an equal distance stays behind the earlier index)
#pragma unroll
                    for (int s = 0; s < KMAX; ++s)
                        if (s == k - 1) {
                            bd[s] = d;
                            bi[s] = base + p;
                        }
#pragma unroll
                    for (int s = KMAX - 1; s > 0; --s)
                        if (s < k && bd[s] < bd[s - 1]) {
                            double td = bd[s]; bd[s] = bd[s - 1]; bd[s - 1] = td;
                            int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
                        }
#pragma unroll
                    for (int s = 0; s < KMAX; ++s)
                        if (s == k - 1) worst = bd[s];
                }
            }
        }
    }
    if (!valid) return;
    int lab[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; ++s) lab[s] = (s < k && bi[s] >= 0) ? __ldg(labels + bi[s]) : 0;
    int best = -1, best_count = 0;
#pragma unroll
    for (int s = 0; s < KMAX; ++s) {
        if (s < k && bi[s] >= 0) {
            int cnt = 0;
#pragma unroll
            for (int u = 0; u <= s; ++u) cnt += (lab[u] == lab[s]) ? 1 : 0;
            if (cnt > best_count) {
                best = lab[s];
                best_count = cnt;
            }
        }
    }
    out_labels[j] = best;
    const bool known = best >= 0 && best < 9;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out_colors[(size_t)j * 3 + c] = known ? kLabelColors[best][c] : (unsigned char)0;
}

static inline int grid_for(long total, int threads) {
    long blocks = ceil_div<long>(total, threads);
    long cap = 148L * 32;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace pn2

using namespace pn2;

PN2_API int pn2_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist,
                         int *idx, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n >= 0 && m > 0);
    if (b == 0 || n == 0) return PN2_OK;
    PN2_REQUIRE_PTR(xyz1);
    PN2_REQUIRE_PTR(xyz2);
    PN2_REQUIRE_PTR(dist);
    PN2_REQUIRE_PTR(idx);
    dim3 grid((unsigned)ceil_div(n, NN_THREADS), (unsigned)b);
    launch_k(three_nn_kernel, grid, NN_THREADS, 0, as_stream(s), n, m, xyz1, xyz2, dist, idx);
    return finish_launch();
}

PN2_API int pn2_fp_weights(int rows, const float *dist, float *weight, pn2_stream_t s) {
    PN2_REQUIRE(rows >= 0);
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dist);
    PN2_REQUIRE_PTR(weight);
    launch_k(fp_weights_kernel, grid_for(rows, 256), 256, 0, as_stream(s), rows, dist, weight);
    return finish_launch();
}

PN2_API int pn2_three_interpolate_ld(int b, int m, int c, int n, const float *points,
                                     const int *idx, const float *weight, float *out, int ldo,
                                     pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && m > 0 && c >= 0 && n >= 0 && ldo >= c);
    if (b == 0 || n == 0 || c == 0) return PN2_OK;
    PN2_REQUIRE_PTR(points);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(weight);
    PN2_REQUIRE_PTR(out);
    const bool vec = (c % 4 == 0) && (ldo % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(points) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    cudaStream_t st = as_stream(s);
    if (vec) {
        long total = (long)b * n * (c / 4);
        launch_k(three_interpolate_kernel<true>, grid_for(total, 256), 256, 0, st, 
            m, c, n, total, points, idx, weight, out, ldo);
    } else {
        long total = (long)b * n * c;
        launch_k(three_interpolate_kernel<false>, grid_for(total, 256), 256, 0, st, 
            m, c, n, total, points, idx, weight, out, ldo);
    }
    return finish_launch();
}

PN2_API int pn2_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                                  const float *weight, float *out, pn2_stream_t s) {
    return pn2_three_interpolate_ld(b, m, c, n, points, idx, weight, out, c, s);
}

PN2_API int pn2_three_interpolate_grad_ld(int b, int n, int c, int m, const float *grad_out,
                                          int ldg, const int *idx, const float *weight,
                                          float *grad_points, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && m > 0 && c >= 0 && n >= 0 && ldg >= c);
    if (b == 0 || c == 0) return PN2_OK;
    PN2_REQUIRE_PTR(grad_points);
    cudaStream_t st = as_stream(s);
    int rc = cuda_status(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, st));
    if (rc) return rc;
    long total = (long)b * n * c;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(grad_out);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(weight);
    if ((c % 4) == 0 && (reinterpret_cast<uintptr_t>(grad_points) & 15) == 0)
        launch_k(three_interpolate_grad_v4_kernel, grid_for(total / 4, 256), 256, 0, st, 
            m, c, n, total / 4, grad_out, ldg, idx, weight, grad_points);
    else
        launch_k(three_interpolate_grad_kernel, grid_for(total, 256), 256, 0, st, 
            m, c, n, total, grad_out, ldg, idx, weight, grad_points);
    return finish_launch();
}

PN2_API int pn2_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out,
                                       const int *idx, const float *weight, float *grad_points,
                                       pn2_stream_t s) {
    return pn2_three_interpolate_grad_ld(b, n, c, m, grad_out, c, idx, weight, grad_points, s);
}

PN2_API int pn2_copy_cols(long rows, int cols, const float *src, int lds, float *dst, int ldd,
                          int accumulate, pn2_stream_t s) {
    PN2_REQUIRE(rows >= 0 && cols >= 0 && lds >= cols && ldd >= cols);
    long total = rows * cols;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(src);
    PN2_REQUIRE_PTR(dst);
    if ((cols % 4) == 0 && (lds % 4) == 0 && (ldd % 4) == 0 && total < (1L << 32) &&
        ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0)
        launch_k(copy_cols_v4_kernel, grid_for(total / 4, 256), 256, 0, as_stream(s), 
            (unsigned)rows, (unsigned)(cols / 4), reinterpret_cast<const float4 *>(src), (unsigned)(lds / 4),
            reinterpret_cast<float4 *>(dst), (unsigned)(ldd / 4), accumulate);
    else
        launch_k(copy_cols_kernel, grid_for(total, 256), 256, 0, as_stream(s), rows, cols, total, src, lds,
                                                                         dst, ldd, accumulate);
    return finish_launch();
}

PN2_API int pn2_interpolate_label_with_color(int num_sparse, int num_dense,
                                             const float *sparse_points, const int *sparse_labels,
                                             const float *dense_points, int *dense_labels,
                                             unsigned char *dense_colors, int knn,
                                             pn2_stream_t s) {
    PN2_REQUIRE(num_sparse >= 0 && num_dense >= 0 && knn > 0);
    if (knn > 32) return PN2_EUNSUPPORTED;  // the candidate list lives in registers
    if (num_dense == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dense_points);
    PN2_REQUIRE_PTR(dense_labels);
    PN2_REQUIRE_PTR(dense_colors);
    if (num_sparse > 0) {
        PN2_REQUIRE_PTR(sparse_points);
        PN2_REQUIRE_PTR(sparse_labels);
    }
    cudaStream_t st = as_stream(s);
    const unsigned grid = (unsigned)ceil_div<long>(num_dense, KV_THREADS);
#define PN2_LAUNCH_VOTE(KM)                                                                    \
    launch_k(knn_vote_kernel<KM>, grid, KV_THREADS, 0, st, num_sparse, num_dense, knn, sparse_points, \
                                                     sparse_labels, dense_points, dense_labels, \
                                                     dense_colors)
    if (knn <= 4) PN2_LAUNCH_VOTE(4);
    else if (knn <= 8) PN2_LAUNCH_VOTE(8);
    else if (knn <= 16) PN2_LAUNCH_VOTE(16);
    else PN2_LAUNCH_VOTE(32);
#undef PN2_LAUNCH_VOTE
    return finish_launch();
}
