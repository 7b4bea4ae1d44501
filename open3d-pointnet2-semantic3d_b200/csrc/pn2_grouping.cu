// pn2_grouping.cu -- ball query, group_point(+grad), fused group+centre+concat (+grad),
// selection sort, for sm_100a.
//
// Replaces tf_ops/tf_grouping.cu of the reference (not a translation):
//
// Ball query.  The reference runs b*256 threads in total and re-reads every data point from
// global memory for every query.  Here a CTA owns 128 queries of one cloud and streams the
// cloud through shared memory in 1024-point tiles staged by the TMA engine
// (cp.async.bulk + mbarrier, double buffered); each thread scans the tile with broadcast
// LDS.128 reads, 4 points per iteration.  When b*m is too small to fill 148 SMs the cloud is
// made fully resident and SPLIT lanes share one query (contiguous chunks, in-order merge).
//
// Bit-exactness: the reference tests  max(sqrtf(d2), 1e-20f) < radius  with IEEE sqrt.  sqrt is
// monotone and correctly rounded, so that predicate equals  !(d2 >= T)  where T is the smallest
// float whose correctly rounded sqrt is >= radius; T is found on the host (pn2_ball_threshold).
// d2 itself is fma(dz,dz,fma(dx,dx,dy*dy)) as in the compiled reference.  The NaN behaviour
// (CUDA max(NaN,1e-20f) = 1e-20f, hence a hit) is preserved by the negated comparison.
#include <math.h>

#include <cub/block/block_scan.cuh>

#include "pn2_common.cuh"

namespace pn2 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

constexpr int BQ_THREADS = 128;
constexpr int BQ_TILE = 1024;  // points per shared-memory stage (12 KB)

// One thread per query, cloud streamed through two TMA-filled stages.
template <bool USE_TMA>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_stream_kernel(int n, int m, float thr, int nsample, const float *__restrict__ xyz1,
                         const float *__restrict__ xyz2, int *__restrict__ idx,
                         int *__restrict__ pts_cnt) {
    pdl_enter();
    __shared__ __align__(128) float tile[2][BQ_TILE * 3];
    __shared__ __align__(8) uint64_t full[2];

    const int cloud = blockIdx.y;
    const int j = blockIdx.x * BQ_THREADS + threadIdx.x;
    const float *data = xyz1 + (size_t)cloud * n * 3;
    const int ntiles = (n + BQ_TILE - 1) / BQ_TILE;

    if (USE_TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&full[0], 1);
            mbar_init(&full[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int t = 0; t < 2 && t < ntiles; ++t) {
                int cntp = min(BQ_TILE, n - t * BQ_TILE);
                mbar_expect_tx(&full[t], cntp * 12);
                tma_load_1d(tile[t], data + (size_t)t * BQ_TILE * 3, cntp * 12, &full[t]);
            }
        }
    }

    const bool valid = j < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float *q = xyz2 + ((size_t)cloud * m + j) * 3;
        qx = __ldg(q);
        qy = __ldg(q + 1);
        qz = __ldg(q + 2);
    }
    int *row = idx + ((size_t)cloud * m + (valid ? j : 0)) * nsample;
    int cnt = 0, first = 0;
    bool done = !valid;

    for (int t = 0; t < ntiles; ++t) {
        const int base = t * BQ_TILE;
        const int cntp = min(BQ_TILE, n - base);
        float *buf = tile[t & 1];
        if (USE_TMA) {
            mbar_wait(&full[t & 1], (t >> 1) & 1);
        } else {
            for (int e = threadIdx.x; e < cntp * 3; e += BQ_THREADS)
                buf[e] = __ldg(data + (size_t)base * 3 + e);
            __syncthreads();
        }
        if (!done) {
            const int k4 = cntp & ~3;
            const float4 *b4 = reinterpret_cast<const float4 *>(buf);
            int k = 0;
            for (; k < k4 && cnt < nsample; k += 4) {
                // 4 points = 12 floats = 3 broadcast LDS.128
                float4 a = b4[(k >> 2) * 3 + 0], bb = b4[(k >> 2) * 3 + 1], c = b4[(k >> 2) * 3 + 2];
                float d0 = sqdist_ref(qx - a.x, qy - a.y, qz - a.z);
                float d1 = sqdist_ref(qx - a.w, qy - bb.x, qz - bb.y);
                float d2 = sqdist_ref(qx - bb.z, qy - bb.w, qz - c.x);
                float d3 = sqdist_ref(qx - c.y, qy - c.z, qz - c.w);
                bool h0 = !(d0 >= thr), h1 = !(d1 >= thr), h2 = !(d2 >= thr), h3 = !(d3 >= thr);
                if (h0 | h1 | h2 | h3) {
                    if (h0 && cnt < nsample) { if (cnt == 0) first = base + k; row[cnt++] = base + k; }
                    if (h1 && cnt < nsample) { if (cnt == 0) first = base + k + 1; row[cnt++] = base + k + 1; }
                    if (h2 && cnt < nsample) { if (cnt == 0) first = base + k + 2; row[cnt++] = base + k + 2; }
                    if (h3 && cnt < nsample) { if (cnt == 0) first = base + k + 3; row[cnt++] = base + k + 3; }
                }
            }
            for (; k < cntp && cnt < nsample; ++k) {
                float d0 = sqdist_ref(qx - buf[k * 3], qy - buf[k * 3 + 1], qz - buf[k * 3 + 2]);
                if (!(d0 >= thr)) {
                    if (cnt == 0) first = base + k;
                    row[cnt++] = base + k;
                }
            }
            done = cnt >= nsample;
        }
        // everyone is finished with this stage; stop early when the whole CTA is done
        const int all_done = __syncthreads_and(done ? 1 : 0);
        if (all_done) {
            // never leave the CTA with a bulk copy still in flight towards its shared memory
            if (USE_TMA && t + 1 < ntiles) mbar_wait(&full[(t + 1) & 1], ((t + 1) >> 1) & 1);
            break;
        }
        if (USE_TMA && threadIdx.x == 0 && t + 2 < ntiles) {
            int c2 = min(BQ_TILE, n - (t + 2) * BQ_TILE);
            mbar_expect_tx(&full[t & 1], c2 * 12);
            tma_load_1d(buf, data + (size_t)(t + 2) * BQ_TILE * 3, c2 * 12, &full[t & 1]);
        }
    }
    if (valid) {
        for (int l = cnt; l < nsample; ++l) row[l] = first;  // pad with the first hit (0 if none)
        pts_cnt[(size_t)cloud * m + j] = cnt;
    }
}

// Small-grid variant: the whole cloud is made resident in shared memory by ONE TMA bulk copy and
// SPLIT lanes share a query.  Per round the SPLIT lanes test SPLIT*4 consecutive points (lane s
// takes points base+4s .. base+4s+3: three LDS.128, bank-conflict free), so lane order == index
// order inside a round and hits can be appended to the output row in place with a SPLIT-lane
// prefix sum.  Rounds without any hit in the warp (the common case) cost one vote instruction.
template <int SPLIT>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_resident_kernel(int n, int m, float thr, int nsample, int use_tma,
                           const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                           int *__restrict__ idx, int *__restrict__ pts_cnt) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int QPB = BQ_THREADS / SPLIT;
    const int npad = ((n + SPLIT * 4 - 1) / (SPLIT * 4)) * (SPLIT * 4);  // whole rounds
    float *pts = reinterpret_cast<float *>(smem_raw);                      // npad*3 floats
    __shared__ __align__(8) uint64_t full;

    const int cloud = blockIdx.y;
    const float *data = xyz1 + (size_t)cloud * n * 3;
    // padding points (never hits: coordinates NaN-free and far away is not needed, they are
    // masked by k < n below); zero them so no uninitialised shared memory is read
    for (int e = n * 3 + threadIdx.x; e < npad * 3; e += BQ_THREADS) pts[e] = 0.f;
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = (uint32_t)n * 12u;
            mbar_expect_tx(&full, total);
            for (uint32_t off = 0; off < total; off += 65536u) {
                uint32_t bytes = min(65536u, total - off);
                tma_load_1d(reinterpret_cast<unsigned char *>(pts) + off,
                            reinterpret_cast<const unsigned char *>(data) + off, bytes, &full);
            }
        }
        mbar_wait(&full, 0);
    } else {
        for (int e = threadIdx.x; e < n * 3; e += BQ_THREADS) pts[e] = __ldg(data + e);
    }
    __syncthreads();

    const int q = threadIdx.x / SPLIT, sub = threadIdx.x % SPLIT;
    const unsigned lane = threadIdx.x & 31;
    const unsigned gbase = lane & ~(unsigned)(SPLIT - 1);
    const int j = blockIdx.x * QPB + q;
    const bool valid = j < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float *qp = xyz2 + ((size_t)cloud * m + j) * 3;
        qx = __ldg(qp);
        qy = __ldg(qp + 1);
        qz = __ldg(qp + 2);
    }
    int *row = idx + ((size_t)cloud * m + (valid ? j : 0)) * nsample;
    int cnt = 0, first = 0;  // uniform inside a SPLIT-lane group
    const float4 *b4 = reinterpret_cast<const float4 *>(pts);
    for (int base = 0; base < npad; base += SPLIT * 4) {
        const int k = base + sub * 4;
        const float4 a = b4[(k >> 2) * 3 + 0], bb = b4[(k >> 2) * 3 + 1], c = b4[(k >> 2) * 3 + 2];
        const float d0 = sqdist_ref(qx - a.x, qy - a.y, qz - a.z);
        const float d1 = sqdist_ref(qx - a.w, qy - bb.x, qz - bb.y);
        const float d2 = sqdist_ref(qx - bb.z, qy - bb.w, qz - c.x);
        const float d3 = sqdist_ref(qx - c.y, qy - c.z, qz - c.w);
        unsigned h = 0;
        if (valid && cnt < nsample) {
            h = (!(d0 >= thr) && k < n ? 1u : 0u) | (!(d1 >= thr) && k + 1 < n ? 2u : 0u) |
                (!(d2 >= thr) && k + 2 < n ? 4u : 0u) | (!(d3 >= thr) && k + 3 < n ? 8u : 0u);
        }
        if (__any_sync(0xFFFFFFFFu, h != 0)) {  // warp-uniform slow path
            const int cme = __popc(h);
            int incl = cme;
#pragma unroll
            for (int off = 1; off < SPLIT; off <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, off, SPLIT);
                if (sub >= off) incl += t;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, SPLIT - 1, SPLIT);
            // first hit of the group (only consulted while cnt == 0)
            const unsigned gm = (__ballot_sync(0xFFFFFFFFu, cme > 0) >> gbase) &
                                (SPLIT == 32 ? 0xFFFFFFFFu : ((1u << SPLIT) - 1u));
            const int src = gm ? (int)gbase + __ffs(gm) - 1 : (int)lane;
            const int f = __shfl_sync(0xFFFFFFFFu, k + __ffs(h) - 1, src);
            if (cnt == 0 && total > 0) first = f;
            int pos = cnt + incl - cme;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if ((h >> i) & 1u) {
                    if (pos < nsample) row[pos] = k + i;
                    ++pos;
                }
            cnt = min(cnt + total, nsample);
            if (__all_sync(0xFFFFFFFFu, !valid || cnt >= nsample)) break;
        }
    }
    if (valid) {
        for (int l = cnt + sub; l < nsample; l += SPLIT) row[l] = first;  // pad (0 if no hit)
        if (sub == 0) pts_cnt[(size_t)cloud * m + j] = cnt;
    }
}

// ---- group_point / grad ------------------------------------------------------------------
template <typename VT>
__global__ void group_point_kernel(int n, int cv, long rows_per_cloud, long total,
                                   const VT *__restrict__ points, const int *__restrict__ idx,
                                   VT *__restrict__ out) {
    pdl_enter();
    // cv = channels in units of VT; e = row*cv + l
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / cv;
        int l = (int)(e - row * cv);
        long cloud = row / rows_per_cloud;
        int ii = __ldg(idx + row);
        out[e] = __ldg(points + (cloud * n + ii) * cv + l);
    }
}

__global__ void group_point_grad_kernel(int n, int c, long rows_per_cloud, long total,
                                        const float *__restrict__ grad_out,
                                        const int *__restrict__ idx,
                                        float *__restrict__ grad_points) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / c;
        int l = (int)(e - row * c);
        long cloud = row / rows_per_cloud;
        int ii = __ldg(idx + row);
        atomicAdd(grad_points + (cloud * n + ii) * c + l, __ldg(grad_out + e));
    }
}

// out[row, :] = [xyz[idx]-new_xyz | points[idx]]  (or [points | xyz] when !xyz_first)
__global__ void group_concat_kernel(int n, int m, int ns, int c, int w, int ld, int xoff, int poff,
                                    int use_xyz, long total, const float *__restrict__ xyz,
                                    const float *__restrict__ new_xyz,
                                    const float *__restrict__ points,
                                    const int *__restrict__ idx, float *__restrict__ out) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / w;
        int col = (int)(e - row * w);
        long grp = row / ns;          // cloud*m + j
        long cloud = grp / m;
        int ii = __ldg(idx + row);
        float v;
        if (use_xyz && col >= xoff && col < xoff + 3) {
            int a = col - xoff;
            v = __fsub_rn(__ldg(xyz + (cloud * n + ii) * 3 + a), __ldg(new_xyz + grp * 3 + a));
        } else {
            v = __ldg(points + (cloud * n + ii) * c + (col - poff));
        }
        out[row * ld + col] = v;
    }
}

// c % 4 == 0: the feature channels of a grouped row as float4 gathers (one thread per row and channel quad, 32-bit
// index arithmetic, the neighbour index read once per quad); the destination columns start at an odd offset
// (3 + 4q behind the centred xyz) so they are written as four scalars, consecutive across the threads of a row.
__global__ void group_concat_feat_v4_kernel(int n, int m_ns, int c, int ld, int poff, unsigned total4,
                                            const float *__restrict__ points, const int *__restrict__ idx,
                                            float *__restrict__ out) {
    pdl_enter();
    const unsigned c4 = (unsigned)c >> 2;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += gridDim.x * blockDim.x) {
        const unsigned row = e / c4, q = e - row * c4;
        const unsigned cloud = row / (unsigned)m_ns;
        const int ii = __ldg(idx + row);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(points + ((size_t)cloud * n + ii) * c) + q);
        float *o = out + (size_t)row * ld + poff + 4 * q;
        o[0] = v.x;
        o[1] = v.y;
        o[2] = v.z;
        o[3] = v.w;
    }
}
// the three centred coordinates of every grouped row
__global__ void group_concat_xyz_kernel(int n, int ns, int m, int ld, int xoff, unsigned rows,
                                        const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                        const int *__restrict__ idx, float *__restrict__ out) {
    pdl_enter();
    for (unsigned row = blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += gridDim.x * blockDim.x) {
        const unsigned grp = row / (unsigned)ns, cloud = grp / (unsigned)m;
        const int ii = __ldg(idx + row);
        const float *p = xyz + ((size_t)cloud * n + ii) * 3, *q = new_xyz + (size_t)grp * 3;
        float *o = out + (size_t)row * ld + xoff;
        o[0] = __fsub_rn(__ldg(p), __ldg(q));
        o[1] = __fsub_rn(__ldg(p + 1), __ldg(q + 1));
        o[2] = __fsub_rn(__ldg(p + 2), __ldg(q + 2));
    }
}

// Feature part of the gradient with vector reductions: one thread per (grouped row, 4 consecutive feature
// channels) -> ONE red.global.add.v4.f32 instead of four scalar atomics (c % 4 == 0 makes the destination
// 16-byte aligned; the source row pitch 3+c is odd, so the four gradients are loaded as scalars).
__global__ void group_concat_grad_feat_v4_kernel(int n, int m, int ns, int c, int w, int poff, long total4,
                                                 const float *__restrict__ grad_out,
                                                 const int *__restrict__ idx,
                                                 float *__restrict__ grad_points) {
    pdl_enter();
    const int c4 = c >> 2;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4;
         e += (long)gridDim.x * blockDim.x) {
        const long row = e / c4;
        const int q = (int)(e - row * c4);
        const long cloud = row / ((long)m * ns);
        const int ii = __ldg(idx + row);
        const float *g = grad_out + row * w + poff + 4 * q;
        float *dst = grad_points + (cloud * n + ii) * c + 4 * q;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__ldg(g)), "f"(__ldg(g + 1)),
                     "f"(__ldg(g + 2)), "f"(__ldg(g + 3))
                     : "memory");
    }
}

__global__ void group_concat_grad_kernel(int n, int m, int ns, int c, int w, int xoff, int poff,
                                         int use_xyz, long total,
                                         const float *__restrict__ grad_out,
                                         const int *__restrict__ idx,
                                         float *__restrict__ grad_points,
                                         float *__restrict__ grad_xyz,
                                         float *__restrict__ grad_new_xyz) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / w;
        int col = (int)(e - row * w);
        long grp = row / ns;
        long cloud = grp / m;
        int ii = __ldg(idx + row);
        float g = __ldg(grad_out + e);
        if (use_xyz && col >= xoff && col < xoff + 3) {
            int a = col - xoff;
            if (grad_xyz) atomicAdd(grad_xyz + (cloud * n + ii) * 3 + a, g);
            if (grad_new_xyz) atomicAdd(grad_new_xyz + grp * 3 + a, -g);
        } else if (grad_points) {
            atomicAdd(grad_points + (cloud * n + ii) * c + (col - poff), g);
        }
    }
}

// ---- selection sort / k nearest neighbours (tf_grouping.cu:95-136, tf_grouping.py:64-89) ----------
// The reference runs k passes of a swap-based selection sort over a row of n distances: pass s finds
// the first minimum (strict '<') of positions s..n-1 of the CURRENT array and swaps it with position
// s.  The swaps permute the array, so among equal distances "first" means first in the permuted
// order, not lowest index.  A warp reproduces that exactly WITHOUT the array: positions whose content
// differs from the original row are at most k (each pass displaces one element), so they are kept in
// a small per-warp table (position, value, original index) plus one bit per position; a pass scans the
// untouched positions straight from the row source (a distance matrix, or distances evaluated on the
// fly from the coordinates -- the fused kNN never materialises the (b,m,n) matrix) and then the table.
// NaN follows the reference's comparisons: `p[t] < p[min]` is false for NaN on either side.
constexpr int SEL_WARPS = 8;
constexpr int SEL_KMAX = 128;  // table entries per warp (one per pass)

struct RowFromMatrix {
    const float *d;
    __device__ __forceinline__ float operator()(int t) const { return __ldg(d + t); }
};
template <int C>
struct RowFromXyz {  // tf_grouping.py:79-82: reduce_sum((xyz1 - xyz2)**2, -1), fp32, left to right
    const float *x1;
    const float *q;
    int c;
    __device__ __forceinline__ float operator()(int t) const {
        const float *p = x1 + (size_t)t * (C > 0 ? C : c);
        float acc = 0.f;
        if (C > 0) {
#pragma unroll
            for (int i = 0; i < (C > 0 ? C : 1); ++i) {
                const float df = __fsub_rn(__ldg(p + i), q[i]);
                acc = i == 0 ? __fmul_rn(df, df) : __fadd_rn(acc, __fmul_rn(df, df));
            }
        } else {
            for (int i = 0; i < c; ++i) {
                const float df = __fsub_rn(__ldg(p + i), __ldg(q + i));
                acc = i == 0 ? __fmul_rn(df, df) : __fadd_rn(acc, __fmul_rn(df, df));
            }
        }
        return acc;
    }
};

// (value, position) of the warp's best candidate under the reference's scan order: smaller value,
// then lower position; NaN never wins.  pos == INT_MAX: no candidate.
__device__ __forceinline__ void sel_reduce(float &v, int &pos, int &aux) {
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const float ov = __shfl_xor_sync(0xFFFFFFFFu, v, off);
        const int op = __shfl_xor_sync(0xFFFFFFFFu, pos, off);
        const int oa = __shfl_xor_sync(0xFFFFFFFFu, aux, off);
        if (op != 0x7FFFFFFF && (pos == 0x7FFFFFFF || ov < v || (ov == v && op < pos))) {
            v = ov;
            pos = op;
            aux = oa;
        }
    }
}

// One row: writes the first min(k,n) selected (value, original index) pairs through `emit(s, v, i)`;
// `bits` (n bits, zeroed by the caller's warp), tpos/tval/tidx: the warp's table in shared memory.
template <typename Row, typename Emit>
__device__ __forceinline__ int selection_passes(const Row &row, int n, int k, unsigned *bits, int *tpos,
                                                float *tval, int *tidx, Emit emit) {
    const int lane = threadIdx.x & 31;
    int nt = 0;
    const int passes = k < n ? k : n;
    for (int s = 0; s < passes; ++s) {
        // content of position s: displaced element or the original
        float cs_v;
        int cs_i;
        {
            float v = 0.f;
            int pos = 0x7FFFFFFF, aux = 0;
            for (int e = lane; e < nt; e += 32)
                if (tpos[e] == s) {
                    v = tval[e];
                    pos = s;
                    aux = tidx[e];
                }
            // at most one live entry per position: a plain "any lane has it" reduction
            const unsigned has = __ballot_sync(0xFFFFFFFFu, pos == s);
            if (has) {
                const int src = __ffs(has) - 1;
                cs_v = __shfl_sync(0xFFFFFFFFu, v, src);
                cs_i = __shfl_sync(0xFFFFFFFFu, aux, src);
            } else {
                cs_v = row(s);
                cs_i = s;
            }
        }
        // untouched positions t > s (original content), then displaced ones
        float bv = 0.f;
        int bp = 0x7FFFFFFF, bi = 0;
        for (int t = s + 1 + lane; t < n; t += 32) {
            if (bits[t >> 5] >> (t & 31) & 1u) continue;
            const float d = row(t);
            if (d == d && (bp == 0x7FFFFFFF || d < bv)) {  // ascending t per lane: strict '<' keeps the first
                bv = d;
                bp = t;
                bi = t;
            }
        }
        for (int e = lane; e < nt; e += 32) {
            const int tp = tpos[e];
            const float d = tval[e];
            if (tp > s && d == d && (bp == 0x7FFFFFFF || d < bv || (d == bv && tp < bp))) {
                bv = d;
                bp = tp;
                bi = tidx[e];
            }
        }
        sel_reduce(bv, bp, bi);
        // the reference keeps min = s unless something is strictly smaller than the content of s
        const bool move = bp != 0x7FFFFFFF && bv < cs_v;
        const float out_v = move ? bv : cs_v;
        const int out_i = move ? bi : cs_i;
        emit(s, out_v, out_i);
        if (move) {  // content of s goes to position bp
            bool updated = false;
            for (int e = lane; e < nt; e += 32)
                if (tpos[e] == bp) {
                    tval[e] = cs_v;
                    tidx[e] = cs_i;
                    updated = true;
                }
            const bool any = __any_sync(0xFFFFFFFFu, updated);
            if (!any && lane == 0) {
                tpos[nt] = bp;
                tval[nt] = cs_v;
                tidx[nt] = cs_i;
                bits[bp >> 5] |= 1u << (bp & 31);
            }
            if (!any) ++nt;
        }
        __syncwarp();
    }
    return nt;
}

__global__ void __launch_bounds__(32 * SEL_WARPS)
selection_sort_kernel(long rows, int n, int k, const float *__restrict__ dist, int *__restrict__ outi,
                      float *__restrict__ out) {
    pdl_enter();
    extern __shared__ unsigned sel_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words = (n + 31) >> 5;
    unsigned *bits = sel_smem + (size_t)warp * words;
    int *tpos = reinterpret_cast<int *>(sel_smem + (size_t)SEL_WARPS * words) + warp * SEL_KMAX;
    float *tval = reinterpret_cast<float *>(sel_smem + (size_t)SEL_WARPS * words + SEL_WARPS * SEL_KMAX) +
                  warp * SEL_KMAX;
    int *tidx = reinterpret_cast<int *>(sel_smem + (size_t)SEL_WARPS * words + 2 * SEL_WARPS * SEL_KMAX) +
                warp * SEL_KMAX;
    for (long row = blockIdx.x * (long)SEL_WARPS + warp; row < rows; row += (long)gridDim.x * SEL_WARPS) {
        for (int w = lane; w < words; w += 32) bits[w] = 0u;
        __syncwarp();
        const float *d = dist + row * n;
        float *o = out + row * n;
        int *oi = outi + row * n;
        RowFromMatrix src{d};
        const int nt = selection_passes(src, n, k, bits, tpos, tval, tidx, [&](int s, float v, int i) {
            if (lane == 0) {
                o[s] = v;
                oi[s] = i;
            }
        });
        // the tail of the permuted array (the reference returns it, tf_grouping.cu:108-113): original
        // content except at displaced positions
        const int passes = k < n ? k : n;
        for (int t = passes + lane; t < n; t += 32) {
            if (!(bits[t >> 5] >> (t & 31) & 1u)) {
                o[t] = __ldg(d + t);
                oi[t] = t;
            }
        }
        for (int e = lane; e < nt; e += 32)
            if (tpos[e] >= passes) {
                o[tpos[e]] = tval[e];
                oi[tpos[e]] = tidx[e];
            }
        __syncwarp();
    }
}

template <int C>
__global__ void __launch_bounds__(32 * SEL_WARPS)
knn_point_kernel(int n, int c, int m, int k, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                 float *__restrict__ val, int *__restrict__ idx) {
    pdl_enter();
    extern __shared__ unsigned sel_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words = (n + 31) >> 5;
    unsigned *bits = sel_smem + (size_t)warp * words;
    int *tpos = reinterpret_cast<int *>(sel_smem + (size_t)SEL_WARPS * words) + warp * SEL_KMAX;
    float *tval = reinterpret_cast<float *>(sel_smem + (size_t)SEL_WARPS * words + SEL_WARPS * SEL_KMAX) +
                  warp * SEL_KMAX;
    int *tidx = reinterpret_cast<int *>(sel_smem + (size_t)SEL_WARPS * words + 2 * SEL_WARPS * SEL_KMAX) +
                warp * SEL_KMAX;
    const int cloud = blockIdx.y;
    const int cc = C > 0 ? C : c;
    const float *x1 = xyz1 + (size_t)cloud * n * cc;
    for (int j = blockIdx.x * SEL_WARPS + warp; j < m; j += gridDim.x * SEL_WARPS) {
        for (int w = lane; w < words; w += 32) bits[w] = 0u;
        __syncwarp();
        const float *q = xyz2 + ((size_t)cloud * m + j) * cc;
        float qreg[C > 0 ? C : 1];
        if (C > 0) {
#pragma unroll
            for (int i = 0; i < (C > 0 ? C : 1); ++i) qreg[i] = __ldg(q + i);
        }
        RowFromXyz<C> src{x1, C > 0 ? qreg : q, cc};
        float *ov = val + ((size_t)cloud * m + j) * k;
        int *oi = idx + ((size_t)cloud * m + j) * k;
        selection_passes(src, n, k, bits, tpos, tval, tidx, [&](int s, float v, int i) {
            if (lane == 0) {
                ov[s] = v;
                oi[s] = i;
            }
        });
        __syncwarp();
    }
}

static inline int grid_for(long total, int threads) {
    long blocks = ceil_div<long>(total, threads);
    long cap = 148L * 32;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace pn2

using namespace pn2;

// ---- ball query over a hashed uniform grid (EXPERIMENTAL: never run on a GPU yet) ----------------
// Brute force tests b*m*n pairs although a ball usually holds a tiny fraction of the cloud (SA1 of
// config 2: 0.1 % of the box).  Here every cloud is binned once into cells of edge `radius`
// (cell = floor(x * (1/radius)) per axis, clamped; cells are hashed into T >= 2n buckets, so no bounding
// box and no host round trip are needed), the points are copied bucket by bucket into a float4 array
// (x, y, z, original index) and a WARP per query visits only the cells that can hold a hit:
//   * coverage: a point that passes the fp32 distance test has |dx| <= radius*(1+4u) per axis, and
//     the cell function is monotone in x, so every hit lies in cell(x_q - rpad) .. cell(x_q + rpad)
//     with rpad = radius*1.0001 and the bounds rounded outwards (3 cells per axis, rarely 4);
//   * a bucket can hold points of other cells (hash collisions): a candidate counts only if its own
//     cell equals the visited cell, which also makes every point count exactly once;
//   * the distance test is the one of the other kernels (same expression, same threshold), so the set of
//     hits is identical; the reference's order (first nsample hits in ascending index) is restored by
//     ranking the collected indices (rank = number of smaller hits; indices are unique);
//   * more than GQ_CAP hits, more than 4 cells along an axis, a non-finite query or any non-finite data point fall back
//     to the ordered brute-force scan inside the same kernel (NaN distances are hits in the reference).
constexpr int GQ_WARPS = 8;
constexpr int GQ_CAP = 512;
constexpr float GQ_CLAMP = 1073741824.0f;  // 2^30

__device__ __forceinline__ int grid_cell(float x, float inv) {
    float v = floorf(__fmul_rn(x, inv));
    return (int)fminf(fmaxf(v, -GQ_CLAMP), GQ_CLAMP);
}
__device__ __forceinline__ unsigned grid_bucket(int cx, int cy, int cz, unsigned mask) {
    return (((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cz * 83492791u)) & mask;
}
__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}

// pass 1: bucket populations (cursor must be zero on entry)
__global__ void grid_count_kernel(int n, long total, unsigned tmask, float inv,
                                  const float *__restrict__ xyz, int *__restrict__ cursor,
                                  int *__restrict__ flag) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const long cloud = e / n;
        const float x = __ldg(xyz + e * 3), y = __ldg(xyz + e * 3 + 1), z = __ldg(xyz + e * 3 + 2);
        if (!finite3(x, y, z)) flag[0] = 1;
        const unsigned bk = grid_bucket(grid_cell(x, inv), grid_cell(y, inv), grid_cell(z, inv), tmask);
        atomicAdd(cursor + cloud * (long)(tmask + 1) + bk, 1);
    }
}

// pass 2: starts = exclusive scan of the populations (one CTA per cloud), cursor = starts
__global__ void __launch_bounds__(1024)
grid_scan_kernel(int T, int *__restrict__ cursor, int *__restrict__ starts, int starts_pitch) {
    pdl_enter();
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    int *cur = cursor + (long)blockIdx.x * T;
    int *st = starts + (long)blockIdx.x * starts_pitch;
    const int per = T / 1024;  // T is a power of two >= 1024
    const int first = threadIdx.x * per;
    int sum = 0;
    for (int i = 0; i < per; ++i) sum += cur[first + i];
    int offset;
    Scan(tmp).ExclusiveSum(sum, offset);
    for (int i = 0; i < per; ++i) {
        const int c = cur[first + i];
        st[first + i] = offset;
        cur[first + i] = offset;
        offset += c;
    }
    if (threadIdx.x == 1023) st[T] = offset;
}

// pass 3: copy every point into its bucket's segment
__global__ void grid_fill_kernel(int n, long total, unsigned tmask, float inv,
                                 const float *__restrict__ xyz, int *__restrict__ cursor,
                                 float4 *__restrict__ sorted) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const long cloud = e / n;
        const int k = (int)(e - cloud * n);
        const float x = __ldg(xyz + e * 3), y = __ldg(xyz + e * 3 + 1), z = __ldg(xyz + e * 3 + 2);
        const unsigned bk = grid_bucket(grid_cell(x, inv), grid_cell(y, inv), grid_cell(z, inv), tmask);
        const int pos = atomicAdd(cursor + cloud * (long)(tmask + 1) + bk, 1);
        sorted[cloud * n + pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

// pass 4: one warp per query
__global__ void __launch_bounds__(GQ_WARPS * 32)
ball_query_grid_kernel(int n, int m, long queries, unsigned tmask, int starts_pitch, float thr,
                       float inv, float rpad, int nsample, const float *__restrict__ xyz1,
                       const float *__restrict__ xyz2, const int *__restrict__ starts,
                       const float4 *__restrict__ sorted, const int *__restrict__ flag,
                       int *__restrict__ idx, int *__restrict__ pts_cnt) {
    pdl_enter();
    __shared__ int hits[GQ_WARPS][GQ_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const long q = (long)blockIdx.x * GQ_WARPS + warp;
    if (q >= queries) return;  // whole warps leave together; no block-wide barrier below
    const long cloud = q / m;
    const float qx = __ldg(xyz2 + q * 3), qy = __ldg(xyz2 + q * 3 + 1), qz = __ldg(xyz2 + q * 3 + 2);
    int *row = idx + q * (long)nsample;
    const int *st = starts + cloud * starts_pitch;
    const float4 *pts = sorted + cloud * n;
    int *my = hits[warp];

    bool brute = __ldg(flag) != 0 || !finite3(qx, qy, qz);
    int lx = 0, ly = 0, lz = 0, hx = 0, hy = 0, hz = 0;
    if (!brute) {
        lx = grid_cell(__fsub_rd(qx, rpad), inv);
        hx = grid_cell(__fadd_ru(qx, rpad), inv);
        ly = grid_cell(__fsub_rd(qy, rpad), inv);
        hy = grid_cell(__fadd_ru(qy, rpad), inv);
        lz = grid_cell(__fsub_rd(qz, rpad), inv);
        hz = grid_cell(__fadd_ru(qz, rpad), inv);
        // per-axis spans (a clamped cell range can be 2^31 wide: never multiply before bounding them)
        const long sx = (long)hx - lx + 1, sy = (long)hy - ly + 1, sz = (long)hz - lz + 1;
        brute = sx > 4 || sy > 4 || sz > 4;
    }
    int h = 0;
    if (!brute) {
        for (int cx = lx; cx <= hx; ++cx)
            for (int cy = ly; cy <= hy; ++cy)
                for (int cz = lz; cz <= hz; ++cz) {
                    const unsigned bk = grid_bucket(cx, cy, cz, tmask);
                    const int s = __ldg(st + bk), e = __ldg(st + bk + 1);
                    for (int base = s; base < e; base += 32) {
                        const int p = base + lane;
                        bool hit = false;
                        int k = 0;
                        if (p < e) {
                            const float4 v = __ldg(pts + p);
                            k = __float_as_int(v.w);
                            hit = grid_cell(v.x, inv) == cx && grid_cell(v.y, inv) == cy &&
                                  grid_cell(v.z, inv) == cz &&
                                  !(sqdist_ref(qx - v.x, qy - v.y, qz - v.z) >= thr);
                        }
                        const unsigned mk = __ballot_sync(0xFFFFFFFFu, hit);
                        if (hit) {
                            const int pos = h + __popc(mk & lt);
                            if (pos < GQ_CAP) my[pos] = k;
                        }
                        h += __popc(mk);
                    }
                }
        brute = h > GQ_CAP;
    }
    int cnt = 0, first = 0;
    if (brute) {
        // ordered scan of the whole cloud: lane order == index order, append in place
        const float *data = xyz1 + cloud * (long)n * 3;
        for (int base = 0; base < n && cnt < nsample; base += 32) {
            const int k = base + lane;
            bool hit = false;
            if (k < n) {
                const float d = sqdist_ref(qx - __ldg(data + k * 3), qy - __ldg(data + k * 3 + 1),
                                           qz - __ldg(data + k * 3 + 2));
                hit = !(d >= thr);
            }
            const unsigned mk = __ballot_sync(0xFFFFFFFFu, hit);
            if (mk) {
                if (cnt == 0) first = base + __ffs(mk) - 1;
                if (hit) {
                    const int pos = cnt + __popc(mk & lt);
                    if (pos < nsample) row[pos] = k;
                }
                cnt += __popc(mk);
            }
        }
        cnt = min(cnt, nsample);
    } else {
        __syncwarp();
        int vmin = 0x7FFFFFFF;
        for (int i = lane; i < h; i += 32) {
            const int v = my[i];
            vmin = min(vmin, v);
            int r = 0;
            for (int u = 0; u < h; ++u) r += my[u] < v ? 1 : 0;
            if (r < nsample) row[r] = v;
        }
        vmin = __reduce_min_sync(0xFFFFFFFFu, vmin);
        cnt = min(h, nsample);
        first = h > 0 ? vmin : 0;
    }
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;  // pad with the first hit (0 if none)
    if (lane == 0) pts_cnt[q] = cnt;
}

static inline int grid_buckets(int n) {
    int t = 1024;
    while (t < 2 * n && t < (1 << 28)) t *= 2;
    return t;
}

// smallest float T >= 0 such that sqrtf(T) >= radius (host sqrtf is IEEE, correctly rounded)
static float ball_threshold(float radius) {
    if (!(radius > 1e-20f)) return 0.f;  // max(d,1e-20f) < radius can never hold
    if (isinf(radius)) return INFINITY;
    float t = (float)((double)radius * (double)radius);
    while (t > 0.f && sqrtf(nextafterf(t, 0.f)) >= radius) t = nextafterf(t, 0.f);
    while (sqrtf(t) < radius) t = nextafterf(t, INFINITY);
    return t;
}

PN2_API float pn2_ball_threshold(float radius) { return ball_threshold(radius); }

PN2_API int pn2_query_ball_point(int b, int n, int m, float radius, int nsample,
                                 const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                                 pn2_stream_t s) {
    // tf_grouping.cpp:80-87: "QueryBallPoint expects positive radius / nsample"
    PN2_REQUIRE(radius > 0.f && nsample > 0);
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    if (b == 0 || m == 0) return PN2_OK;
    PN2_REQUIRE_PTR(xyz1);
    PN2_REQUIRE_PTR(xyz2);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(pts_cnt);
    cudaStream_t st = as_stream(s);
    const float thr = ball_threshold(radius);
    // TMA bulk copies need 16-byte aligned sources and sizes: every cloud starts at n*12 bytes
    const bool tma_ok = (reinterpret_cast<uintptr_t>(xyz1) % 16 == 0) && (n % 4 == 0);

    // small grids: make the cloud resident and split each query over SPLIT lanes
    const long queries = (long)b * m;
    int split = 1;
    while (split < 8 && queries * split < 148L * 1024) split *= 2;
    const size_t res_smem = (size_t)(((n + 8 * 4 - 1) / (8 * 4)) * (8 * 4)) * 12;  // padded to whole rounds
    if (split > 1 && res_smem <= 200 * 1024) {
        dim3 grid((unsigned)ceil_div(m, BQ_THREADS / split), (unsigned)b);
        int rc = PN2_OK;
#define PN2_LAUNCH_RES(SP)                                                                      \
    do {                                                                                        \
        auto kern = ball_query_resident_kernel<SP>;                                             \
        rc = opt_in_dyn_smem(kern, res_smem);                                                   \
        if (rc) return rc;                                                                      \
        launch_k(kern, grid, BQ_THREADS, res_smem, st, n, m, thr, nsample, tma_ok ? 1 : 0, xyz1,     \
                                                 xyz2, idx, pts_cnt);                           \
    } while (0)
        if (split == 2) PN2_LAUNCH_RES(2);
        else if (split == 4) PN2_LAUNCH_RES(4);
        else PN2_LAUNCH_RES(8);
#undef PN2_LAUNCH_RES
        return finish_launch();
    }
    dim3 grid((unsigned)ceil_div(m, BQ_THREADS), (unsigned)b);
    if (tma_ok)
        launch_k(ball_query_stream_kernel<true>, grid, BQ_THREADS, 0, st, n, m, thr, nsample, xyz1, xyz2,
                                                                    idx, pts_cnt);
    else
        launch_k(ball_query_stream_kernel<false>, grid, BQ_THREADS, 0, st, n, m, thr, nsample, xyz1,
                                                                     xyz2, idx, pts_cnt);
    return finish_launch();
}

PN2_API int pn2_group_point(int b, int n, int c, int m, int nsample, const float *points,
                            const int *idx, float *out, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && nsample >= 0);
    long rows = (long)b * m * nsample;
    if (rows == 0 || c == 0) return PN2_OK;
    PN2_REQUIRE_PTR(points);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(out);
    cudaStream_t st = as_stream(s);
    const bool vec = (c % 4 == 0) && (reinterpret_cast<uintptr_t>(points) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    if (vec) {
        long total = rows * (c / 4);
        launch_k(group_point_kernel<float4>, grid_for(total, 256), 256, 0, st, 
            n, c / 4, (long)m * nsample, total, reinterpret_cast<const float4 *>(points), idx,
            reinterpret_cast<float4 *>(out));
    } else {
        long total = rows * c;
        launch_k(group_point_kernel<float>, grid_for(total, 256), 256, 0, st, 
            n, c, (long)m * nsample, total, points, idx, out);
    }
    return finish_launch();
}

PN2_API int pn2_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out,
                                 const int *idx, float *grad_points, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && nsample >= 0);
    if (b == 0 || c == 0) return PN2_OK;
    PN2_REQUIRE_PTR(grad_points);
    cudaStream_t st = as_stream(s);
    int rc = cuda_status(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, st));
    if (rc) return rc;
    long total = (long)b * m * nsample * c;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(grad_out);
    PN2_REQUIRE_PTR(idx);
    launch_k(group_point_grad_kernel, grid_for(total, 256), 256, 0, st, n, c, (long)m * nsample, total,
                                                                  grad_out, idx, grad_points);
    return finish_launch();
}

PN2_API int pn2_group_concat_ld(int b, int n, int m, int nsample, int c, const float *xyz,
                                const float *new_xyz, const float *points, const int *idx,
                                int xyz_first, int use_xyz, float *out, int ld, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample >= 0 && c >= 0);
    PN2_REQUIRE(use_xyz || c > 0);
    const int w = (use_xyz ? 3 : 0) + c;
    PN2_REQUIRE(ld >= w);
    long total = (long)b * m * nsample * w;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(out);
    if (use_xyz) {
        PN2_REQUIRE_PTR(xyz);
        PN2_REQUIRE_PTR(new_xyz);
    }
    if (c > 0) PN2_REQUIRE_PTR(points);
    const int xoff = xyz_first ? 0 : c, poff = (use_xyz && xyz_first) ? 3 : 0;
    const long rows = (long)b * m * nsample;
    if (c > 0 && (c % 4) == 0 && (reinterpret_cast<uintptr_t>(points) & 15) == 0 && rows * (c / 4) < (1L << 32) &&
        rows < (1L << 31)) {
        cudaStream_t st = as_stream(s);
        launch_k(group_concat_feat_v4_kernel, grid_for(rows * (c / 4), 256), 256, 0, st, 
            n, m * nsample, c, ld, poff, (unsigned)(rows * (c / 4)), points, idx, out);
        int rc = finish_launch();
        if (rc || !use_xyz) return rc;
        launch_k(group_concat_xyz_kernel, grid_for(rows, 256), 256, 0, st, n, nsample, m, ld, xoff, (unsigned)rows, xyz,
                                                                    new_xyz, idx, out);
        return finish_launch();
    }
    launch_k(group_concat_kernel, grid_for(total, 256), 256, 0, as_stream(s), 
        n, m, nsample, c, w, ld, xoff, poff, use_xyz, total, xyz, new_xyz, points, idx, out);
    return finish_launch();
}

PN2_API int pn2_group_concat(int b, int n, int m, int nsample, int c, const float *xyz,
                             const float *new_xyz, const float *points, const int *idx,
                             int xyz_first, int use_xyz, float *out, pn2_stream_t s) {
    return pn2_group_concat_ld(b, n, m, nsample, c, xyz, new_xyz, points, idx, xyz_first, use_xyz,
                               out, (use_xyz ? 3 : 0) + c, s);
}

PN2_API int pn2_group_concat_grad(int b, int n, int m, int nsample, int c, const float *grad_out,
                                  const int *idx, int xyz_first, int use_xyz, float *grad_points,
                                  float *grad_xyz, float *grad_new_xyz, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample >= 0 && c >= 0);
    if (b == 0) return PN2_OK;
    cudaStream_t st = as_stream(s);
    int rc;
    if (grad_points && c > 0) {
        rc = cuda_status(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, st));
        if (rc) return rc;
    }
    if (grad_xyz) {
        rc = cuda_status(cudaMemsetAsync(grad_xyz, 0, sizeof(float) * (size_t)b * n * 3, st));
        if (rc) return rc;
    }
    if (grad_new_xyz) {
        rc = cuda_status(cudaMemsetAsync(grad_new_xyz, 0, sizeof(float) * (size_t)b * m * 3, st));
        if (rc) return rc;
    }
    const int w = (use_xyz ? 3 : 0) + c;
    long total = (long)b * m * nsample * w;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(grad_out);
    PN2_REQUIRE_PTR(idx);
    const int xoff = xyz_first ? 0 : c, poff = (use_xyz && xyz_first) ? 3 : 0;
    if (grad_points && c > 0 && (c % 4) == 0 && (reinterpret_cast<uintptr_t>(grad_points) & 15) == 0) {
        // feature channels by vector reductions; the xyz columns keep the scalar kernel (and are skipped
        // altogether when nobody asked for xyz gradients: the model never does)
        const long total4 = (long)b * m * nsample * (c / 4);
        launch_k(group_concat_grad_feat_v4_kernel, grid_for(total4, 256), 256, 0, st, n, m, nsample, c, w, poff, total4,
                                                                                grad_out, idx, grad_points);
        rc = finish_launch();
        if (rc) return rc;
        if (!(use_xyz && (grad_xyz || grad_new_xyz))) return PN2_OK;
        grad_points = nullptr;  // done
    }
    launch_k(group_concat_grad_kernel, grid_for(total, 256), 256, 0, st, 
        n, m, nsample, c, w, xoff, poff, use_xyz, total, grad_out, idx, grad_points, grad_xyz,
        grad_new_xyz);
    return finish_launch();
}

static size_t selection_smem(int n) {
    return ((size_t)SEL_WARPS * ((n + 31) / 32) + 3 * (size_t)SEL_WARPS * SEL_KMAX) * 4;
}

PN2_API int pn2_selection_sort(int b, int n, int m, int k, const float *dist, int *outi,
                               float *out, pn2_stream_t s) {
    PN2_REQUIRE(k > 0);  // tf_grouping.cpp:142-144 "SelectionSort expects positive k"
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    long rows = (long)b * m;
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dist);
    PN2_REQUIRE_PTR(outi);
    PN2_REQUIRE_PTR(out);
    if (k > SEL_KMAX && n > SEL_KMAX) return PN2_EUNSUPPORTED;  // table of displaced elements
    const size_t smem = selection_smem(n);
    if (smem > 200 * 1024) return PN2_EUNSUPPORTED;
    int rc = opt_in_dyn_smem(selection_sort_kernel, smem);
    if (rc) return rc;
    long blocks = ceil_div<long>(rows, SEL_WARPS);
    if (blocks > 148L * 8) blocks = 148L * 8;
    launch_k(selection_sort_kernel, (unsigned)blocks, 32 * SEL_WARPS, smem, as_stream(s), rows, n, k, dist, outi, out);
    return finish_launch();
}

/* Fused kNN (tf_grouping.py:64-89 without its (b,m,n) distance tensor): val (b,m,k) squared distances,
 * idx (b,m,k), the first k entries of the reference's selection sort, bit for bit (ties included). */
PN2_API int pn2_knn_point(int b, int n, int c, int m, int k, const float *xyz1, const float *xyz2,
                          float *val, int *idx, pn2_stream_t s) {
    PN2_REQUIRE(k > 0 && c > 0);
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    PN2_REQUIRE(k <= n);  // tf.slice(outi, [0,0,0], [-1,-1,k]) needs k <= n
    if (b == 0 || m == 0) return PN2_OK;
    PN2_REQUIRE_PTR(xyz1);
    PN2_REQUIRE_PTR(xyz2);
    PN2_REQUIRE_PTR(val);
    PN2_REQUIRE_PTR(idx);
    if (k > SEL_KMAX) return PN2_EUNSUPPORTED;
    const size_t smem = selection_smem(n);
    if (smem > 200 * 1024) return PN2_EUNSUPPORTED;
    unsigned gx = (unsigned)ceil_div(m, SEL_WARPS);
    if ((long)gx * b > 148L * 16) gx = (unsigned)max(1L, 148L * 16 / b);
    dim3 grid(gx, (unsigned)b);
    int rc;
    if (c == 3) {
        rc = opt_in_dyn_smem(knn_point_kernel<3>, smem);
        if (rc) return rc;
        launch_k(knn_point_kernel<3>, grid, 32 * SEL_WARPS, smem, as_stream(s), n, c, m, k, xyz1, xyz2, val, idx);
    } else {
        rc = opt_in_dyn_smem(knn_point_kernel<0>, smem);
        if (rc) return rc;
        launch_k(knn_point_kernel<0>, grid, 32 * SEL_WARPS, smem, as_stream(s), n, c, m, k, xyz1, xyz2, val, idx);
    }
    return finish_launch();
}

PN2_API long pn2_ball_grid_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 16;
    const long T = grid_buckets(n);
    return 16 + (long)b * ((T + 4) * 4 + T * 4 + (long)n * 16);
}

PN2_API int pn2_query_ball_point_grid(int b, int n, int m, float radius, int nsample,
                                      const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                                      void *workspace, long workspace_bytes, pn2_stream_t s) {
    PN2_REQUIRE(radius > 0.f && nsample > 0);
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    if (b == 0 || m == 0) return PN2_OK;
    PN2_REQUIRE_PTR(xyz1);
    PN2_REQUIRE_PTR(xyz2);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(pts_cnt);
    PN2_REQUIRE_PTR(workspace);
    PN2_REQUIRE(workspace_bytes >= pn2_ball_grid_workspace_bytes(b, n));
    PN2_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0);
    cudaStream_t st = as_stream(s);
    const float thr = ball_threshold(radius);
    const int T = grid_buckets(n);
    const int pitch = T + 4;
    // layout: [flag: 4 ints][sorted: b*n float4][starts: b*(T+4) ints][cursor: b*T ints]
    int *flag = static_cast<int *>(workspace);
    float4 *sorted = reinterpret_cast<float4 *>(flag + 4);
    int *starts = reinterpret_cast<int *>(sorted + (size_t)b * n);
    int *cursor = starts + (size_t)b * pitch;
    int rc = cuda_status(cudaMemsetAsync(flag, 0, 16, st));
    if (rc) return rc;
    rc = cuda_status(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)b * T, st));
    if (rc) return rc;
    const float inv = 1.0f / radius;
    const float rpad = nextafterf(radius * 1.0001f, INFINITY);
    const long total = (long)b * n;
    long blocks = ceil_div<long>(total, 256);
    if (blocks > 148L * 16) blocks = 148L * 16;
    launch_k(grid_count_kernel, (int)blocks, 256, 0, st, n, total, (unsigned)(T - 1), inv, xyz1, cursor, flag);
    rc = finish_launch();
    if (rc) return rc;
    launch_k(grid_scan_kernel, b, 1024, 0, st, T, cursor, starts, pitch);
    rc = finish_launch();
    if (rc) return rc;
    launch_k(grid_fill_kernel, (int)blocks, 256, 0, st, n, total, (unsigned)(T - 1), inv, xyz1, cursor, sorted);
    rc = finish_launch();
    if (rc) return rc;
    const long queries = (long)b * m;
    const long qblocks = ceil_div<long>(queries, GQ_WARPS);
    PN2_REQUIRE(qblocks < (1l << 31));
    launch_k(ball_query_grid_kernel, (unsigned)qblocks, GQ_WARPS * 32, 0, st, 
        n, m, queries, (unsigned)(T - 1), pitch, thr, inv, rpad, nsample, xyz1, xyz2, starts, sorted, flag,
        idx, pts_cnt);
    return finish_launch();
}
