// pn2_gemm_tc.cu -- shared-MLP GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   Y[M,N] = f(A)[M,K] * Bt[N,K]^T (+ bias) (+ column statistics)       (forward and dgrad)
//
// Precision: the reference computes these 1x1 convolutions in fp32 and the parity bar is 1e-5
// absolute, which plain TF32 (10-bit mantissa) cannot meet.  Every product is therefore
// evaluated as an error-compensated 3xTF32 sum with fp32 accumulation in tensor memory:
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   a_hi = rna_tf32(a), a_lo = a - a_hi
// (the dropped a_lo*b_lo term is ~2^-22 relative).  Effective tensor peak = TF32 peak / 3.
// The tensor core adds into its fp32 accumulator with truncation, a biased error that grows
// with the number of accumulations (measured: 3.5e-5 at K=768 with one accumulator), so the
// large hi*hi products and the small correction products go to TWO accumulators (main / corr)
// that are summed with a round-to-nearest add in the epilogue; K is limited to 512.
//
// Structure (one persistent CTA per SM, 18 warps, roles as in the canonical Blackwell GEMM):
//   warps 0-3  epilogue: tcgen05.ld accumulator (lane quadrant = warp), transpose through
//              padded shared memory, + bias, coalesced row-major stores, fp64 column statistics
//   warps 4-15 A producers, three independent groups of 4 warps that take K chunks round-robin
//              (so three chunks of global loads are in flight per SM: the kernel was load-latency
//              bound with one): coalesced float4 loads of a 128 x 32 chunk, previous layer's
//              BatchNorm affine + ReLU applied on the fly, hi/lo split, st.shared into the
//              128B-swizzled K-major UMMA layout, fence.proxy.async, mbarrier arrive
//   warp 16    B loader: one cp.async.bulk per K chunk from a pre-split, pre-swizzled weight
//              image in global memory (built by tc_prep_b_kernel, L2 resident)
//   warp 17    TMEM allocation + single-thread tcgen05.mma issue (3 MMAs per K=8 step),
//              tcgen05.commit onto the "stage empty" / "accumulator full" mbarriers
// The accumulator pair is double buffered in TMEM (2 x 2 x N columns, N <= 128 per pass) so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>

#include "pn2_common.cuh"

namespace pn2 {
namespace tc {

constexpr int BM = 128;       // rows per tile (UMMA M)
constexpr int BK = 32;        // fp32/tf32 elements per K chunk = one 128-byte swizzle row
constexpr int NPG = 2;                    // independent A-producer groups (chunks of loads in flight)
constexpr int THREADS = 32 * (4 + 4 * NPG + 2);  // 4 epilogue + 4*NPG producer + loader + MMA warps
constexpr int W_LOADER = 4 + 4 * NPG, W_MMA = 5 + 4 * NPG;
constexpr int MAX_STAGES = 4;
constexpr int EPI_LD = 36;    // padded row length (floats) of the epilogue transpose buffer

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B -> 64), [46,48) version = 1,
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}
// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32, fp32 accumulate,
// both operands K-major: c_format=F32 [4,6), a/b_format=TF32 [7,10)/[10,13), N>>3 [17,23),
// M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of element (row, k) inside one K-major SWIZZLE_128B chunk image (row = 128 B)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) +
                      (k & 3) * 4);
}

// Weight image: for every K chunk kc: [hi image: Npad rows x 128 B][lo image: same], elements
// Bt(n,k) = src[n*s_n + k*s_k], zero outside (N,K).
__global__ void tc_prep_b_kernel(int N, int K, int Npad, int KC, const float *__restrict__ src,
                                 long s_n, long s_k, float *__restrict__ image) {
    const long total = (long)KC * Npad * BK;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(e % BK);
        const long t = e / BK;
        const int n = (int)(t % Npad);
        const int kc = (int)(t / Npad);
        const int k = kc * BK + kk;
        float v = (n < N && k < K) ? __ldg(src + n * s_n + k * s_k) : 0.f;
        const float hi = tf32_rna(v);
        const float lo = v - hi;
        unsigned char *base = reinterpret_cast<unsigned char *>(image) + (size_t)kc * 2 * Npad * 128;
        const uint32_t off = sw128_offset(n, kk);
        *reinterpret_cast<float *>(base + off) = hi;
        *reinterpret_cast<float *>(base + (size_t)Npad * 128 + off) = lo;
    }
}

// Cycle accounting of CTA 0 (one thread per role), read back by pn2_debug_tc_trace():
//  [0] mma: cycles waiting for a full stage   [1] mma: cycles issuing   [2] mma: waiting acc_empty
//  [3] producer g0: load+transform            [4] producer g0: waiting empty   [5] producer g0: store
//  [6] epilogue w0: waiting acc_full          [7] epilogue w0: processing      [8] loader: waiting empty
//  [9] total kernel cycles (mma thread)       [10] chunks                      [11] tiles
__device__ long long g_tc_trace[16];

struct Params {
    long M;
    int K, N, Npad, KC, stages, lda, ldy, a_relu;
    const float *A, *a_scale, *a_shift, *bias, *image;
    float *Y;
    double *stats_sum, *stats_sq;  // per-column sum / sum of squares (fp64), or NULL
};

__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve: stages x [A_hi 16K | A_lo 16K | B_hi Npad*128 | B_lo Npad*128], epilogue staging, barriers
    const uint32_t a_bytes = BM * 128;
    const uint32_t b_bytes = (uint32_t)p.Npad * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    unsigned char *stage_base = smem;
    float *epi = reinterpret_cast<float *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi + 4 * 32 * EPI_LD);
    uint64_t *full = bars;                    // [stages]  A producers (128) + B loader (1) + tx
    uint64_t *empty = bars + MAX_STAGES;      // [stages]  tcgen05.commit
    uint64_t *acc_full = bars + 2 * MAX_STAGES;       // [2]
    uint64_t *acc_empty = bars + 2 * MAX_STAGES + 2;  // [2] epilogue threads (128)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Nacc = (p.Npad + 31) & ~31;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(4 * Nacc)) ncols <<= 1;  // {main, corr} x double buffer

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 129);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const long num_tiles = (p.M + BM - 1) / BM;

    if (warp >= 4 && warp < W_LOADER) {
        // ================================ A producers ================================
        // group g takes chunks g, g+npg, ... of this CTA's (tile, kc) sequence; npg <= stages
        // keeps every group within one ring revolution of the consumer (phase parity is safe)
        const int g = (warp - 4) >> 2;
        const int npg = NPG < p.stages ? NPG : p.stages;
        const int t = (threadIdx.x - 128) & 127;
        const int k4 = t & 7, r0 = t >> 3;  // float4 slot inside the 32-wide chunk, base row
        const bool vec_ok = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
        const long my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const long total_chunks = my_tiles * p.KC;
        if (g < npg) {
            for (long itl = g; itl < total_chunks; itl += npg) {
                const uint32_t it = (uint32_t)itl;
                const long tile = blockIdx.x + (itl / p.KC) * gridDim.x;
                const int kc = (int)(itl % p.KC);
                const long m0 = tile * BM;
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                const int kbase = kc * BK + k4 * 4;
                const bool tr = (blockIdx.x == 0 && g == 0 && t == 0);
                const long long c0 = tr ? clock64() : 0;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long m = m0 + r0 + 16 * i;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (m < p.M) {
                        const float *src = p.A + m * p.lda + kbase;
                        if (vec_ok && kbase + 3 < p.K) {
                            x = __ldg(reinterpret_cast<const float4 *>(src));
                        } else {
                            if (kbase + 0 < p.K) x.x = __ldg(src + 0);
                            if (kbase + 1 < p.K) x.y = __ldg(src + 1);
                            if (kbase + 2 < p.K) x.z = __ldg(src + 2);
                            if (kbase + 3 < p.K) x.w = __ldg(src + 3);
                        }
                    }
                    v[i] = x;
                }
                if (p.a_scale) {
                    float sc[4], sh[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = kbase + j < p.K;
                        sc[j] = ok ? __ldg(p.a_scale + kbase + j) : 0.f;
                        sh[j] = ok ? __ldg(p.a_shift + kbase + j) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool row_ok = (m0 + r0 + 16 * i) < p.M;
                        float *e = reinterpret_cast<float *>(&v[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float y = __fmaf_rn(e[j], sc[j], sh[j]);
                            if (p.a_relu) y = fmaxf(y, 0.f);
                            e[j] = row_ok ? y : 0.f;
                        }
                    }
                }
                // keep the transformed values live up to here so c1 really is "data has arrived"
                const long long c1 = tr ? (clock64() + (long long)(__float_as_int(v[7].w) & 0)) : 0;
                mbar_wait(&empty[s], ph ^ 1);
                const long long c2 = tr ? clock64() : 0;
                unsigned char *a_hi = stage_base + (size_t)s * stage_bytes;
                unsigned char *a_lo = a_hi + a_bytes;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = r0 + 16 * i;
                    const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 +
                                                    (((k4 ^ (r & 7)) & 7) << 4));
                    float4 hi, lo;
                    hi.x = tf32_rna(v[i].x); lo.x = v[i].x - hi.x;
                    hi.y = tf32_rna(v[i].y); lo.y = v[i].y - hi.y;
                    hi.z = tf32_rna(v[i].z); lo.z = v[i].z - hi.z;
                    hi.w = tf32_rna(v[i].w); lo.w = v[i].w - hi.w;
                    *reinterpret_cast<float4 *>(a_hi + off) = hi;
                    *reinterpret_cast<float4 *>(a_lo + off) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&full[s]);
                if (tr) {
                    const long long c3 = clock64();
                    g_tc_trace[3] += c1 - c0;
                    g_tc_trace[4] += c2 - c1;
                    g_tc_trace[5] += c3 - c2;
                }
            }
        }
    } else if (warp == W_LOADER) {
        // ================================ B loader ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < p.KC; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    const long long c0 = blockIdx.x == 0 ? clock64() : 0;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (blockIdx.x == 0) g_tc_trace[8] += clock64() - c0;
                    unsigned char *b_hi = stage_base + (size_t)s * stage_bytes + 2 * a_bytes;
                    mbar_expect_tx(&full[s], 2 * b_bytes);
                    bulk_g2s(b_hi,
                             reinterpret_cast<const unsigned char *>(p.image) + (size_t)kc * 2 * b_bytes,
                             2 * b_bytes, &full[s]);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.Npad);
            uint32_t it = 0, tcnt = 0;
            const bool tr = blockIdx.x == 0;
            const long long k0c = tr ? clock64() : 0;
            long long w_full = 0, w_issue = 0, w_acc = 0;
            for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcnt) {
                const uint32_t acc = tcnt & 1, aph = (tcnt >> 1) & 1;
                const long long ca = tr ? clock64() : 0;
                mbar_wait(&acc_empty[acc], aph ^ 1);
                if (tr) w_acc += clock64() - ca;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * (uint32_t)(2 * Nacc);  // main
                const uint32_t dc = d + (uint32_t)Nacc;                     // corrections
                for (int kc = 0; kc < p.KC; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    const long long cw = tr ? clock64() : 0;
                    mbar_wait(&full[s], ph);
                    const long long ci = tr ? clock64() : 0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(stage_base + (size_t)s * stage_bytes);
                    const uint64_t dah = make_desc(a_hi), dal = make_desc(a_hi + a_bytes);
                    const uint64_t dbh = make_desc(a_hi + 2 * a_bytes),
                                   dbl = make_desc(a_hi + 2 * a_bytes + b_bytes);
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2);  // 32 bytes per K=8 step, >>4
                        const uint32_t accum = (kc > 0 || kk > 0) ? 1u : 0u;
                        umma_tf32(d, dah + adv, dbh + adv, idesc, accum);
                        umma_tf32(dc, dal + adv, dbh + adv, idesc, accum);
                        umma_tf32(dc, dah + adv, dbl + adv, idesc, 1u);
                    }
                    umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
                    if (tr) {
                        w_full += ci - cw;
                        w_issue += clock64() - ci;
                    }
                }
                umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
            }
            if (tr) {
                g_tc_trace[0] += w_full;
                g_tc_trace[1] += w_issue;
                g_tc_trace[2] += w_acc;
                g_tc_trace[9] += clock64() - k0c;
                g_tc_trace[10] += it;
                g_tc_trace[11] += tcnt;
            }
        }
    } else {
        // ================================ epilogue (warps 0-3) ================================
        float *stg = epi + warp * 32 * EPI_LD;
        const int nblk = (p.N + 31) / 32;
        double ssum[4], ssq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ssum[i] = ssq[i] = 0.0;
        uint32_t tcnt = 0;
        for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcnt) {
            const uint32_t acc = tcnt & 1, aph = (tcnt >> 1) & 1;
            const long m0 = tile * BM + warp * 32;
            const bool tr = (blockIdx.x == 0 && threadIdx.x == 0);
            const long long ce0 = tr ? clock64() : 0;
            mbar_wait(&acc_full[acc], aph);
            const long long ce1 = tr ? clock64() : 0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                if (cb >= nblk) break;
                const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                    acc * (uint32_t)(2 * Nacc) + cb * 32;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[16], rc[16];
                    tmem_ld16_nowait(ta + half * 16, r);
                    tmem_ld16_nowait(ta + (uint32_t)Nacc + half * 16, rc);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 o;  // main + corrections, round-to-nearest
                        o.x = __uint_as_float(r[q * 4]) + __uint_as_float(rc[q * 4]);
                        o.y = __uint_as_float(r[q * 4 + 1]) + __uint_as_float(rc[q * 4 + 1]);
                        o.z = __uint_as_float(r[q * 4 + 2]) + __uint_as_float(rc[q * 4 + 2]);
                        o.w = __uint_as_float(r[q * 4 + 3]) + __uint_as_float(rc[q * 4 + 3]);
                        *reinterpret_cast<float4 *>(stg + lane * EPI_LD + half * 16 + q * 4) = o;
                    }
                }
                __syncwarp();
                const int col = cb * 32 + lane;
                const bool col_ok = col < p.N;
                const float bv = (p.bias && col_ok) ? __ldg(p.bias + col) : 0.f;
                // all 32 shared-memory reads are issued back to back (independent), then the stores
                float vals[32];
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) vals[rr] = stg[rr * EPI_LD + lane] + bv;
                __syncwarp();
                const long rows_left = p.M - m0;
                if (col_ok && rows_left > 0) {
                    float *yp = p.Y + m0 * p.ldy + col;
                    if (rows_left >= 32) {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) yp[(long)rr * p.ldy] = vals[rr];
                    } else {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr)
                            if (rr < rows_left) yp[(long)rr * p.ldy] = vals[rr];
                    }
                    if (p.stats_sum) {
                        // shifted fp32 partials (deviations from the first row of the block carry
                        // no cancellation), recombined exactly in fp64
                        const int nv = rows_left >= 32 ? 32 : (int)rows_left;
                        const float c0 = vals[0];
                        float p1 = 0.f, p2 = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr)
                            if (rr < nv) {
                                const float dv = vals[rr] - c0;
                                p1 += dv;
                                p2 = __fmaf_rn(dv, dv, p2);
                            }
                        const double dc = (double)c0, dn = (double)nv, d1 = (double)p1;
                        ssum[cb] += d1 + dn * dc;
                        ssq[cb] += (double)p2 + 2.0 * dc * d1 + dn * dc * dc;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[acc]);
            if (tr) {
                g_tc_trace[6] += ce1 - ce0;
                g_tc_trace[7] += clock64() - ce1;
            }
        }
        if (p.stats_sum) {
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                const int col = cb * 32 + lane;
                if (cb < nblk && col < p.N) {
                    atomicAdd(p.stats_sum + col, ssum[cb]);
                    atomicAdd(p.stats_sq + col, ssq[cb]);
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(ncols)
                     : "memory");
    }
}

static int opt_in_smem(const void *kernel, int slot) {
    static bool done[2][64] = {{false}};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!done[slot][dev]) {
        int rc = cuda_status(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  227 * 1024));
        if (rc) return rc;
        done[slot][dev] = true;
    }
    return PN2_OK;
}

static size_t image_bytes(int K, int N) {
    const int Npad = (N + 15) & ~15, KC = (K + BK - 1) / BK;
    return (size_t)KC * 2 * Npad * 128;
}

// Y[M, n0:n0+Nc] for one chunk of at most 128 output columns (two accumulators x 2 buffers)
static int run_chunk(long M, int K, int Nc, const float *A, int lda, const float *a_scale,
                     const float *a_shift, int a_relu, const float *bsrc, long s_n, long s_k,
                     const float *bias, float *Y, int ldy, double *stats_sum, double *stats_sq,
                     float *ws, cudaStream_t st) {
    Params p;
    p.M = M;
    p.K = K;
    p.N = Nc;
    p.Npad = (Nc + 15) & ~15;
    p.KC = (K + BK - 1) / BK;
    p.lda = lda;
    p.ldy = ldy;
    p.a_relu = a_relu;
    p.A = A;
    p.a_scale = a_scale;
    p.a_shift = a_shift;
    p.bias = bias;
    p.image = ws;
    p.Y = Y;
    p.stats_sum = stats_sum;
    p.stats_sq = stats_sq;
    const size_t stage_bytes = 2 * (size_t)BM * 128 + 2 * (size_t)p.Npad * 128;
    const size_t fixed = 4 * 32 * EPI_LD * sizeof(float) + (2 * MAX_STAGES + 4) * 8 + 16;
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);  // 227 KB usable shared memory per CTA
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 1) return PN2_EUNSUPPORTED;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed;

    const long total = (long)p.KC * p.Npad * BK;
    int pb = (int)((total + 255) / 256);
    if (pb > 148 * 8) pb = 148 * 8;
    tc_prep_b_kernel<<<pb, 256, 0, st>>>(Nc, K, p.Npad, p.KC, bsrc, s_n, s_k, ws);
    int rc = finish_launch();
    if (rc) return rc;

    // opt in to the full 227 KB once per device (not per launch: no driver call on the hot path,
    // nothing that could disturb a stream capture)
    rc = opt_in_smem(reinterpret_cast<const void *>(tc_gemm_kernel), 0);
    if (rc) return rc;
    const long tiles = (M + BM - 1) / BM;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    tc_gemm_kernel<<<grid, THREADS, smem, st>>>(p);
    return finish_launch();
}

}  // namespace tc

// =====================================================================================================
// wgrad on the tensor cores:  dW[K,N] += f(X)[M,K]^T * dY[M,N]   (contraction over the M rows)
//
// Both operands are "MN-major" for the MMA: X^T has its M' = feature index contiguous in memory
// (a row of X), dY^T likewise.  For 32-bit operands the only MN-major shared-memory layout the
// tensor core accepts is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major tf32
// operands, SW128_32B is the only available smem layout"): atoms of 4 contraction rows x 128 B
// (32 fp32 of the feature/column index), the four 32-byte chunks of a row XOR-permuted by the row
// index (Swizzle<2,5,2>).  An atom is 4 consecutive rows of the row-major source, so the
// producers copy rows (transform + hi/lo split) without any transpose.
// The M rows are cut into segments of 512 rows; each segment accumulates into its own TMEM
// accumulator pair (double buffered) and is flushed to dW with fp32 atomics by the epilogue warps.
// Short segments bound the number of truncating tensor-core accumulations (64 per segment).
// =====================================================================================================
namespace tcw {
using namespace tc;

constexpr int W_NPG = 2;                        // independent producer groups (stages of loads in flight)
constexpr int W_WMMA = 4 + 4 * W_NPG;            // MMA / TMEM warp
constexpr int W_THREADS = 32 * (W_WMMA + 1);  // warps 0-3 epilogue, 4.. producers, last MMA
constexpr int W_ROWS = 32;      // contraction rows per stage (4 MMAs of K=8)
constexpr int W_SEG = 512;      // contraction rows per accumulator segment
constexpr int W_FEAT = 128;     // features (rows of dW) per CTA pass = UMMA M

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes,
                                                 uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of float4 slot c4 (MN index 4*c4..4*c4+3) of contraction row r inside a stage
// operand laid out as [k-group r/4][MN-group c4/8] atoms of 512 B
__device__ __forceinline__ uint32_t mn_offset(int r, int c4, int groups) {
    return (uint32_t)(((r >> 2) * groups + (c4 >> 3)) * 512 + (r & 3) * 128 +
                      (((((c4 & 7) >> 1) ^ (r & 3)) & 3) << 5) + ((c4 & 1) << 4));
}

struct WParams {
    long M;
    int K, N, Npad, NG, lda, ldy, ldw, a_relu, k0, n0, stages;
    int dbg;  // PN2_DBG_WGRAD bit mask (diagnostics only): 1 skip atomics, 2 skip dY loads, 4 skip X loads
    const float *A, *a_scale, *a_shift, *dY;
    float *dW;
};

__global__ void __launch_bounds__(W_THREADS, 1) tc_wgrad_kernel(const WParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t a_bytes = 8 * 4 * 512;                  // [k-group 8][mg 4] atoms of 512 B
    const uint32_t b_bytes = 8 * (uint32_t)p.NG * 512;     // [k-group 8][ng NG] atoms
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    float *epi = reinterpret_cast<float *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi + 4 * 32 * EPI_LD);
    uint64_t *full = bars, *empty = bars + MAX_STAGES;
    uint64_t *acc_full = bars + 2 * MAX_STAGES, *acc_empty = bars + 2 * MAX_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Nacc = p.Npad;  // multiple of 32
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(4 * Nacc)) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 128);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_WMMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const long num_seg = (p.M + W_SEG - 1) / W_SEG;

    if (warp >= 4 && warp < W_WMMA) {
        // ================================ producers ================================
        // three independent groups of 128 threads take stages round-robin (three stages of global
        // loads in flight); every thread issues ALL its loads of a stage (8 float4 of X rows,
        // up to 8 float4 of dY rows) before the first use, then splits and stores
        const int g = (warp - 4) >> 2;
        const int npg = W_NPG < p.stages ? W_NPG : p.stages;
        const int tt = (threadIdx.x - 128) & 127;
        const bool a_vec = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) &&
                           (p.k0 % 4 == 0);
        const bool b_vec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dY) & 15) == 0) &&
                           (p.n0 % 4 == 0);
        const int bw4 = p.Npad >> 2;               // float4 per dY row (8, 16, 24 or 32)
        const int b_iters = (W_ROWS * bw4) >> 7;   // Npad/16 <= 8
        const int c4 = tt & 31, rr = tt >> 5;      // A part: float4 slot / base row
        const int kf = p.k0 + 4 * c4;
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.a_scale) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (kf + j < p.K) {
                    sc[j] = __ldg(p.a_scale + kf + j);
                    sh[j] = __ldg(p.a_shift + kf + j);
                }
        }
        // flat stage sequence of this CTA: segment-major, 16 stages per full segment
        const long my_segs = blockIdx.x < num_seg ? (num_seg - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        long itl = 0;
        if (g < npg) {
            for (long sg = 0; sg < my_segs; ++sg) {
                const long seg0 = (blockIdx.x + sg * gridDim.x) * (long)W_SEG;
                const int nst = (int)((min((long)W_SEG, p.M - seg0) + W_ROWS - 1) / W_ROWS);
                for (int sidx = 0; sidx < nst; ++sidx, ++itl) {
                    if ((int)(itl % npg) != g) continue;
                    const uint32_t it = (uint32_t)itl;
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    const long mbase = seg0 + (long)sidx * W_ROWS;
                    unsigned char *st_base = smem + (size_t)s * stage_bytes;
                    float4 va[8], vb[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long m = mbase + rr + 4 * i;
                        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (m < p.M && kf < p.K && !(p.dbg & 4)) {
                            const float *src = p.A + m * p.lda + kf;
                            if (a_vec && kf + 3 < p.K) {
                                x = __ldg(reinterpret_cast<const float4 *>(src));
                            } else {
                                x.x = __ldg(src);
                                if (kf + 1 < p.K) x.y = __ldg(src + 1);
                                if (kf + 2 < p.K) x.z = __ldg(src + 2);
                                if (kf + 3 < p.K) x.w = __ldg(src + 3);
                            }
                        }
                        va[i] = x;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (q < b_iters) {
                            const int e = tt + 128 * q;
                            const int bc4 = e % bw4, r = e / bw4;
                            const long m = mbase + r;
                            const int nf = p.n0 + 4 * bc4;
                            if (m < p.M && nf < p.N && !(p.dbg & 2)) {
                                const float *src = p.dY + m * p.ldy + nf;
                                if (b_vec && nf + 3 < p.N) {
                                    x = __ldg(reinterpret_cast<const float4 *>(src));
                                } else {
                                    x.x = __ldg(src);
                                    if (nf + 1 < p.N) x.y = __ldg(src + 1);
                                    if (nf + 2 < p.N) x.z = __ldg(src + 2);
                                    if (nf + 3 < p.N) x.w = __ldg(src + 3);
                                }
                            }
                        }
                        vb[q] = x;
                    }
                    if (p.a_scale) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool row_ok = (mbase + rr + 4 * i) < p.M;
                            float *e = reinterpret_cast<float *>(&va[i]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float y = __fmaf_rn(e[j], sc[j], sh[j]);
                                if (p.a_relu) y = fmaxf(y, 0.f);
                                e[j] = (row_ok && kf + j < p.K) ? y : 0.f;
                            }
                        }
                    }
                    mbar_wait(&empty[s], ph ^ 1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t off = mn_offset(rr + 4 * i, c4, 4);
                        float4 hi, lo;
                        hi.x = tf32_rna(va[i].x); lo.x = va[i].x - hi.x;
                        hi.y = tf32_rna(va[i].y); lo.y = va[i].y - hi.y;
                        hi.z = tf32_rna(va[i].z); lo.z = va[i].z - hi.z;
                        hi.w = tf32_rna(va[i].w); lo.w = va[i].w - hi.w;
                        *reinterpret_cast<float4 *>(st_base + off) = hi;
                        *reinterpret_cast<float4 *>(st_base + a_bytes + off) = lo;
                    }
                    unsigned char *b_hi = st_base + 2 * a_bytes;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q < b_iters) {
                            const int e = tt + 128 * q;
                            const uint32_t off = mn_offset(e / bw4, e % bw4, p.NG);
                            float4 hi, lo;
                            hi.x = tf32_rna(vb[q].x); lo.x = vb[q].x - hi.x;
                            hi.y = tf32_rna(vb[q].y); lo.y = vb[q].y - hi.y;
                            hi.z = tf32_rna(vb[q].z); lo.z = vb[q].z - hi.z;
                            hi.w = tf32_rna(vb[q].w); lo.w = vb[q].w - hi.w;
                            *reinterpret_cast<float4 *>(b_hi + off) = hi;
                            *reinterpret_cast<float4 *>(b_hi + b_bytes + off) = lo;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&full[s]);
                }
            }
        }
    } else if (warp == W_WMMA) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.Npad) | (1u << 15) | (1u << 16);  // A, B MN-major
            uint32_t it = 0, tcnt = 0;
            const bool tr = blockIdx.x == 0;
            const long long k0c = tr ? clock64() : 0;
            long long w_full = 0, w_issue = 0, w_acc = 0;
            for (long seg = blockIdx.x; seg < num_seg; seg += gridDim.x, ++tcnt) {
                const long seg0 = seg * W_SEG;
                const int nst = (int)((min((long)W_SEG, p.M - seg0) + W_ROWS - 1) / W_ROWS);
                const uint32_t acc = tcnt & 1, aph = (tcnt >> 1) & 1;
                const long long ca = tr ? clock64() : 0;
                mbar_wait(&acc_empty[acc], aph ^ 1);
                if (tr) w_acc += clock64() - ca;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * (uint32_t)(2 * Nacc);
                const uint32_t dc = d + (uint32_t)Nacc;
                for (int sidx = 0; sidx < nst; ++sidx, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    const long long cw = tr ? clock64() : 0;
                    mbar_wait(&full[s], ph);
                    const long long ci = tr ? clock64() : 0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t b_hi = a_hi + 2 * a_bytes;
#pragma unroll
                    for (int kg = 0; kg < 4; ++kg) {
                        // one K=8 MMA spans two 4-row k-groups: LBO = next MN group (512 B),
                        // SBO = next k-group (groups * 512 B)
                        const uint64_t dah = make_desc_mn(a_hi + kg * 4096, 512, 2048);
                        const uint64_t dal = make_desc_mn(a_hi + a_bytes + kg * 4096, 512, 2048);
                        const uint64_t dbh = make_desc_mn(b_hi + kg * p.NG * 1024, 512, p.NG * 512);
                        const uint64_t dbl = make_desc_mn(b_hi + b_bytes + kg * p.NG * 1024, 512, p.NG * 512);
                        const uint32_t accum = (sidx > 0 || kg > 0) ? 1u : 0u;
                        umma_tf32(d, dah, dbh, idesc, accum);
                        umma_tf32(dc, dal, dbh, idesc, accum);
                        umma_tf32(dc, dah, dbl, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (tr) {
                        w_full += ci - cw;
                        w_issue += clock64() - ci;
                    }
                }
                umma_commit(&acc_full[acc]);
            }
            if (tr) {
                g_tc_trace[0] += w_full;
                g_tc_trace[1] += w_issue;
                g_tc_trace[2] += w_acc;
                g_tc_trace[9] += clock64() - k0c;
                g_tc_trace[10] += it;
                g_tc_trace[11] += tcnt;
            }
        }
    } else if (warp < 4) {
        // ================================ epilogue ================================
        float *stg = epi + warp * 32 * EPI_LD;
        const int nblk = p.Npad / 32;
        uint32_t tcnt = 0;
        for (long seg = blockIdx.x; seg < num_seg; seg += gridDim.x, ++tcnt) {
            const uint32_t acc = tcnt & 1, aph = (tcnt >> 1) & 1;
            mbar_wait(&acc_full[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int cb = 0; cb < nblk; ++cb) {
                const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                    acc * (uint32_t)(2 * Nacc) + cb * 32;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[16], rc[16];
                    tmem_ld16_nowait(ta + half * 16, r);
                    tmem_ld16_nowait(ta + (uint32_t)Nacc + half * 16, rc);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 o;
                        o.x = __uint_as_float(r[q * 4]) + __uint_as_float(rc[q * 4]);
                        o.y = __uint_as_float(r[q * 4 + 1]) + __uint_as_float(rc[q * 4 + 1]);
                        o.z = __uint_as_float(r[q * 4 + 2]) + __uint_as_float(rc[q * 4 + 2]);
                        o.w = __uint_as_float(r[q * 4 + 3]) + __uint_as_float(rc[q * 4 + 3]);
                        *reinterpret_cast<float4 *>(stg + lane * EPI_LD + half * 16 + q * 4) = o;
                    }
                }
                __syncwarp();
                float vals[32];
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) vals[rr] = stg[rr * EPI_LD + lane];
                const int n = p.n0 + cb * 32 + lane;
                const int kb = p.k0 + warp * 32;
                if (n < p.N && kb < p.K && !(p.dbg & 1)) {
                    float *wp = p.dW + (long)kb * p.ldw + n;
                    const int kv = p.K - kb;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr)
                        if (rr < kv) atomicAdd(wp + (long)rr * p.ldw, vals[rr]);
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[acc]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_WMMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(ncols)
                     : "memory");
    }
}

static int run(long M, int K, int Kdo, int N, const float *A, int lda, const float *a_scale,
               const float *a_shift, int a_relu, const float *dY, float *dW, cudaStream_t st) {
    for (int n0 = 0; n0 < N; n0 += 128) {
        for (int k0 = 0; k0 < Kdo; k0 += W_FEAT) {
            WParams p;
            p.M = M;
            p.K = K;
            p.N = N;
            const int nc = (N - n0) < 128 ? (N - n0) : 128;
            p.Npad = (nc + 31) & ~31;
            p.NG = p.Npad / 32;
            p.lda = lda;
            p.ldy = N;
            p.ldw = N;
            p.a_relu = a_relu;
            p.k0 = k0;
            p.n0 = n0;
            p.A = A;
            p.a_scale = a_scale;
            p.a_shift = a_shift;
            p.dY = dY;
            p.dW = dW;
            {
                const char *e = getenv("PN2_DBG_WGRAD");
                p.dbg = e ? atoi(e) : 0;
            }
            const size_t stage_bytes = 2 * 16384 + 2 * (size_t)p.NG * 4096;
            const size_t fixed = 4 * 32 * EPI_LD * sizeof(float) + (2 * MAX_STAGES + 4) * 8 + 16;
            int stages = (int)((227 * 1024 - fixed) / stage_bytes);
            if (stages > MAX_STAGES) stages = MAX_STAGES;
            if (stages < 1) return PN2_EUNSUPPORTED;
            p.stages = stages;
            const size_t smem = (size_t)stages * stage_bytes + fixed;
            int rc = tc::opt_in_smem(reinterpret_cast<const void *>(tc_wgrad_kernel), 1);
            if (rc) return rc;
            const long segs = (M + W_SEG - 1) / W_SEG;
            const int grid = (int)(segs < num_sms() ? segs : num_sms());
            tc_wgrad_kernel<<<grid, W_THREADS, smem, st>>>(p);
            rc = finish_launch();
            if (rc) return rc;
        }
    }
    return PN2_OK;
}

}  // namespace tcw

int tc_linear_wgrad(long M, int K, int N, const float *A, int lda, const float *a_scale,
                    const float *a_shift, int a_relu, const float *dY, float *dW, bool force,
                    int *k_done, cudaStream_t st) {
    *k_done = 0;
    if (M < 2048 || N < 16 || K < 4) return PN2_EUNSUPPORTED;
    // measured on B200 (tests/bench_gemm.py): the 128-feature MMA tile only pays off for wide
    // layers with many rows; narrow or short problems are faster on the split-K fp32 kernel
    if (!force && (M < 65536 || K < 64 || N < 64)) return PN2_EUNSUPPORTED;
    // a narrow tail of features (K = 131 -> 3) would cost a full extra pass over dY on the
    // 128-feature MMA tile; it is left to the caller's fp32 kernel (k_done tells where it starts)
    int Kdo = K;
    if (!force && K > tcw::W_FEAT && (K % tcw::W_FEAT) < 32) Kdo = K - (K % tcw::W_FEAT);
    *k_done = Kdo;
    return tcw::run(M, K, Kdo, N, A, lda, a_scale, a_shift, a_relu, dY, dW, st);
}

// Shapes worth the tensor cores: at least one full tile of rows, K and N not tiny.
// K <= 512 keeps the truncating tensor-core accumulation inside the 1e-5 parity bar.
static bool tc_shape_ok(long M, int K, int N) { return M >= 128 && K >= 16 && K <= 512 && N >= 16; }

constexpr int TC_NCHUNK = 128;
static size_t tc_image_bytes(int K, int N) { return tc::image_bytes(K, N > TC_NCHUNK ? TC_NCHUNK : N); }

// one buffer serves both orientations of a layer: forward (K x N) and dgrad (N x K)
size_t tc_workspace_bytes(int K, int N) {
    const size_t a = tc_image_bytes(K, N), b = tc_image_bytes(N, K);
    return a > b ? a : b;
}

int tc_linear_fwd(long M, int K, int N, const float *A, int lda, const float *a_scale,
                  const float *a_shift, int a_relu, const float *W, const float *bias, float *Y,
                  double *stats, float *ws, size_t ws_bytes, cudaStream_t st) {
    if (!tc_shape_ok(M, K, N) || ws == nullptr || ws_bytes < tc_image_bytes(K, N))
        return PN2_EUNSUPPORTED;
    for (int n0 = 0; n0 < N; n0 += TC_NCHUNK) {
        const int nc = (N - n0) < TC_NCHUNK ? (N - n0) : TC_NCHUNK;
        // Bt(n,k) = W[k*N + n0 + n]
        int rc = tc::run_chunk(M, K, nc, A, lda, a_scale, a_shift, a_relu, W + n0, 1, N,
                               bias ? bias + n0 : nullptr, Y + n0, N, stats ? stats + n0 : nullptr,
                               stats ? stats + N + n0 : nullptr, ws, st);
        if (rc) return rc;
    }
    return PN2_OK;
}

int tc_linear_dgrad(long M, int K, int N, const float *dY, const float *W, float *dX, int ldx,
                    float *ws, size_t ws_bytes, cudaStream_t st) {
    // dX[M,K] = dY[M,N] * W[K,N]^T : contraction over N, output columns = K ; Bt(k,n) = W[k*N + n]
    if (!tc_shape_ok(M, N, K) || ws == nullptr || ws_bytes < tc_image_bytes(N, K))
        return PN2_EUNSUPPORTED;
    for (int k0 = 0; k0 < K; k0 += TC_NCHUNK) {
        const int kc = (K - k0) < TC_NCHUNK ? (K - k0) : TC_NCHUNK;
        int rc = tc::run_chunk(M, N, kc, dY, N, nullptr, nullptr, 0, W + (long)k0 * N, N, 1, nullptr,
                               dX + k0, ldx, nullptr, nullptr, ws, st);
        if (rc) return rc;
    }
    return PN2_OK;
}

}  // namespace pn2

// debug hook (not in the public header): read and reset the cycle accounting of CTA 0
extern "C" __attribute__((visibility("default"))) int pn2_debug_tc_trace(long long *out16) {
    long long zeros[16] = {0};
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, pn2::tc::g_tc_trace, sizeof(zeros)) != cudaSuccess) return -2;
    if (cudaMemcpyToSymbol(pn2::tc::g_tc_trace, zeros, sizeof(zeros)) != cudaSuccess) return -2;
    return 0;
}
