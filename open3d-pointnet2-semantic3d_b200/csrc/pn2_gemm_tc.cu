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
// Structure (one persistent CTA per SM, 15 warps, roles as in the canonical Blackwell GEMM):
//   warps 0-7   epilogue: tcgen05.ld (lane quadrant = warp & 3, column blocks split between the two
//               warps of a quadrant) + bias -> 128B-swizzled staging tile -> TMA tensor store;
//               BatchNorm column statistics from the staging tile (fp32 partials, fp64 totals)
//   warps 8-11  A transform: raw 128 x 32 chunk from the TMA ring -> previous layer's BatchNorm
//               affine + ReLU -> hi/lo split -> st.shared into the 128B-swizzled K-major UMMA
//               layout -> fence.proxy.async -> mbarrier arrive
//   warp 12     A loader: 2-D TMA tensor loads into a 4-slot raw ring (zero fill out of range)
//   warp 13     B loader: cp.async.bulk of a pre-split, pre-swizzled weight image in global memory
//               (built by tc_prep_b_kernel, L2 resident); resident in shared memory when it fits
//   warp 14     TMEM allocation + single-thread tcgen05.mma issue (3 MMAs per K=8 step),
//               tcgen05.commit onto the "stage empty" / "accumulator full" mbarriers
// The accumulator pair is double buffered in TMEM (2 x 2 x N columns, N <= 128 per column block) so
// the epilogue of tile i overlaps the main loop of tile i+1; all column blocks of a layer run in one
// launch (CTA c owns block c mod nblocks).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "pn2_common.cuh"

namespace pn2 {
namespace tc {

constexpr int BM = 128;       // rows per tile (UMMA M)
constexpr int BK = 32;        // fp32/tf32 elements per K chunk = one 128-byte swizzle row
constexpr int W_EPI = 8;                  // epilogue warps 0-7: two per TMEM lane quadrant, column blocks split
#ifndef PN2_XF_WARPS
#define PN2_XF_WARPS 8
#endif
constexpr int W_XF = PN2_XF_WARPS;        // A-transform warps: groups of 4 on alternating chunks (A/B builds: 4)
constexpr int THREADS = 32 * (W_EPI + W_XF + 3);  // + A loader + B loader + MMA
constexpr int W_ALOAD = W_EPI + W_XF, W_BLOAD = W_ALOAD + 1, W_MMA = W_ALOAD + 2;
constexpr int MAX_RAW = 4;                // raw A ring slots (16 KB each) filled by TMA tensor loads
constexpr int RAW_BYTES = BM * BK * 4;
constexpr int EPI_BYTES = 32 * 1024;      // epilogue staging
constexpr int MAX_STAGES = 4;
constexpr int MAX_KC = 32;    // K <= 1024 (K > 512 runs as two K halves with their own accumulators)

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
// one lane of a converged warp (CUTLASS elect_one_sync): the MMA warp runs its loop with ALL lanes so that
// descriptors and addresses stay warp-uniform (uniform registers feed UTCHMMA directly); only the elected
// lane issues.  Computing them inside an `if (lane == 0)` region made the compiler move every operand
// through a R2UR "waterfall" loop: ~85 cycles per MMA, the whole issue budget of a chunk.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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

// Weight image: for every column block (128 columns, or all N <= 128) and every K chunk kc:
// [hi image: Npad rows x 128 B][lo image: same], elements Bt(n,k) = src[n*s_n + k*s_k], zero outside (N,K).
__global__ void tc_prep_b_kernel(int N, int Ntot, int K, int Npad, int KC, int nchunks,
                                 const float *__restrict__ src, long s_n, long s_k,
                                 float *__restrict__ image) {
    pdl_enter();
    const long total = (long)nchunks * KC * Npad * BK;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(e % BK);
        const long t = e / BK;
        const int nl = (int)(t % Npad);
        const long t2 = t / Npad;
        const int kc = (int)(t2 % KC);
        const int ch = (int)(t2 / KC);
        const int k = kc * BK + kk;
        const long n = (long)ch * 128 + nl;  // N = columns per block when nchunks > 1
        float v = (nl < N && n < Ntot && k < K) ? __ldg(src + n * s_n + k * s_k) : 0.f;
        const float hi = tf32_rna(v);
        const float lo = v - hi;
        unsigned char *base = reinterpret_cast<unsigned char *>(image) + ((size_t)ch * KC + kc) * 2 * Npad * 128;
        const int n_ = nl;
        const uint32_t off = sw128_offset(n_, kk);
        *reinterpret_cast<float *>(base + off) = hi;
        *reinterpret_cast<float *>(base + (size_t)Npad * 128 + off) = lo;
    }
}

// All weight images of a step in ONE launch: blockIdx.y = table entry, blockIdx.x strides over its
// elements (same element mapping as tc_prep_b_kernel).
__global__ void tc_prep_images_kernel(const pn2_linear_image *__restrict__ table) {
    pdl_enter();
    const pn2_linear_image d = table[blockIdx.y];
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < d.total;
         e += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(e % BK);
        const long t = e / BK;
        const int nl = (int)(t % d.Npad);
        const long t2 = t / d.Npad;
        const int kc = (int)(t2 % d.KC);
        const int ch = (int)(t2 / d.KC);
        const int k = kc * BK + kk;
        const long n = (long)ch * 128 + nl;
        float v = (nl < d.N && n < d.Ntot && k < d.K) ? __ldg(d.src + n * d.s_n + k * d.s_k) : 0.f;
        const float hi = tf32_rna(v);
        const float lo = v - hi;
        unsigned char *base = reinterpret_cast<unsigned char *>(d.image) + ((size_t)ch * d.KC + kc) * 2 * d.Npad * 128;
        const uint32_t off = sw128_offset(nl, kk);
        *reinterpret_cast<float *>(base + off) = hi;
        *reinterpret_cast<float *>(base + (size_t)d.Npad * 128 + off) = lo;
    }
}

// Cycle accounting of CTA 0 (one thread per role), read back by pn2_debug_tc_trace():
//  [0] mma: cycles waiting for a full stage   [1] mma: cycles issuing   [2] mma: waiting acc_empty
//  [3] transform g0: wait raw + load          [4] transform g0: waiting empty  [5] transform g0: store
//  [6] epilogue w0: waiting acc_full          [7] epilogue w0: processing      [8] B loader: waiting empty
//  [9] total kernel cycles (mma thread)       [10] chunks                      [11] tiles
#ifdef PN2_TRACE
__device__ long long g_tc_trace[16];
#define TR_ON(cond) (cond)
#define TR_CLOCK(cond) ((cond) ? clock64() : 0ll)
#define TR_ADD(i, v) (g_tc_trace[i] += (v))
#else  // production build: no instrumentation in the role loops
#define TR_ON(cond) false
#define TR_CLOCK(cond) 0ll
#define TR_ADD(i, v) ((void)(v))
#endif

struct Params {
    long M;
    int K, N, Npad, KC, stages, b_res, raw_slots, a_tma, y_tma, lda, ldy, a_relu;
    int nchunks;  // column blocks of 128 handled by this launch (CTA c works on block c % nchunks)
    int Ntot;     // output columns of the launch; the last block may be narrower than N
    int ksplit;   // K > 512: chunks [0,KC/2) and [KC/2,KC) accumulate separately (no double buffering)
    int stack;    // [B_hi ; B_lo] read as ONE 2*Npad-row operand: 2 MMAs per K=8 step instead of 3
    int epi_alt;  // output at most 32 columns wide: the two epilogue warp groups take ALTERNATE TILES (each
                  // owns one accumulator buffer) instead of alternate column blocks of the same tile
    const float *A, *a_scale, *a_shift, *bias, *image;
    float *Y;
    double *stats_sum, *stats_sq;  // per-column sum / sum of squares (fp64), or NULL
    // fused train-mode BatchNorm finalize (fin_scale != NULL): the LAST CTA to finish turns the column
    // statistics into scale / shift / saved mean+rstd and updates the moving statistics -- what
    // pn2_bn_train_finalize does in a launch of its own (22 launches per training step)
    const float *fin_gamma, *fin_beta;
    float *fin_mm, *fin_mv, *fin_scale, *fin_shift, *fin_saved;
    unsigned *fin_done;  // zeroed by the caller
    float fin_eps, fin_decay;
    int fin_unbiased;
};

__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *tm, int c0, int c1,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
        "%3}], [%4];" ::"r"(smem_u32(dst_smem)),
        "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int c0, int c1,
                                             const void *src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm),
                 "r"(c0), "r"(c1), "r"(smem_u32(src_smem))
                 : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmY,
                   const Params p) {
    pdl_trigger();
    if (threadIdx.x == 0) {  // the descriptors are kernel parameters: fetched while the previous kernel drains
        if (p.a_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        if (p.y_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    }
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve:  [resident B: KC x (B_hi | B_lo)]            (p.b_res: weights loaded once per CTA)
    //         stages x [A_hi 16K | A_lo 16K (| B_hi | B_lo when B is streamed per chunk)]
    //         raw_slots x 16K   raw fp32 A chunks as written by the TMA tensor loads
    //         32K epilogue staging (4 warps x 2 buffers x 32 rows x 128 B, 128B-swizzled)
    //         bias (128 floats), mbarriers
    const uint32_t a_bytes = BM * 128;
    const uint32_t b_bytes = (uint32_t)p.Npad * 128;
    const uint32_t stage_bytes = p.b_res ? 2 * a_bytes : 2 * a_bytes + 2 * b_bytes;
    unsigned char *bres = smem;
    unsigned char *stage_base = smem + (p.b_res ? (size_t)p.KC * 2 * b_bytes : 0);
    unsigned char *raw = stage_base + (size_t)p.stages * stage_bytes;
    unsigned char *epi_b = raw + (size_t)p.raw_slots * RAW_BYTES;
    float *sbias = reinterpret_cast<float *>(epi_b + EPI_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sbias + 128);
    uint64_t *full = bars;                    // [stages]  transform threads (128) (+ B loader + tx)
    uint64_t *empty = bars + MAX_STAGES;      // [stages]  tcgen05.commit
    uint64_t *acc_full = bars + 2 * MAX_STAGES;       // [2]
    uint64_t *acc_empty = bars + 2 * MAX_STAGES + 2;  // [2] epilogue threads (128)
    uint64_t *bfull = bars + 2 * MAX_STAGES + 4;      // [MAX_KC] resident-B chunk landed (tx)
    uint64_t *raw_full = bfull + MAX_KC;              // [MAX_RAW] TMA tx
    uint64_t *raw_empty = raw_full + MAX_RAW;         // [MAX_RAW] transform threads (128)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(raw_empty + MAX_RAW);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Nacc = (p.Npad + 31) & ~31;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(4 * Nacc)) ncols <<= 1;  // {main, corr} x (double buffer | K halves)
    // column block of this CTA and its row-tile sequence
    const int nc = blockIdx.x % p.nchunks, n0 = nc * 128;
    const int Nv = (p.Ntot - n0) < p.N ? (p.Ntot - n0) : p.N;  // valid columns of this block
    const long mt0 = blockIdx.x / p.nchunks, mstride = gridDim.x / p.nchunks;
    const unsigned char *image = reinterpret_cast<const unsigned char *>(p.image) +
                                 (size_t)nc * p.KC * 2 * ((size_t)p.Npad * 128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], p.b_res ? 128 : 129);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], p.epi_alt ? 16 * W_EPI : 32 * W_EPI);
        }
        for (int kc = 0; kc < MAX_KC; ++kc) mbar_init(&bfull[kc], 1);
        for (int r = 0; r < MAX_RAW; ++r) {
            mbar_init(&raw_full[r], 1);
            mbar_init(&raw_empty[r], 128);
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
    // barrier init and TMEM allocation above overlap the tail of the previous kernel; nothing before this line
    // touches global memory (the tensor maps are kernel parameters)
    pdl_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const long num_tiles = (p.M + BM - 1) / BM;

    if (warp >= W_EPI && warp < W_ALOAD) {
        // ================================ A transform ================================
        // The raw chunk comes from the TMA ring (or, when A cannot be described by a tensor map,
        // from float4 global loads); BatchNorm affine + ReLU of the previous layer, hi/lo split,
        // st.shared into the 128B-swizzled K-major UMMA layout, proxy fence, mbarrier arrive.
        // two groups of 128 threads work on alternating chunks (group g: chunks g, g+2, ...): the transform
        // of one chunk is a ~1.8 k-cycle dependent chain (raw wait, ld.shared, affine, split, st.shared,
        // proxy fence), twice the time the tensor core needs for the chunk's 12 MMAs -- two chunks in
        // flight hide it (profiles/README_r02.md, role-cycle trace)
        const int grp = (warp - W_EPI) >> 2;
        const int t = threadIdx.x - 32 * W_EPI - 128 * grp;
        const int k4 = t & 7, r0 = t >> 3;  // float4 slot inside the 32-wide chunk, base row
        const bool vec_ok = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
        const long my_tiles = mt0 < num_tiles ? (num_tiles - mt0 + mstride - 1) / mstride : 0;
        const long total_chunks = my_tiles * p.KC;
        {
            for (long itl = grp; itl < total_chunks; itl += W_XF / 4) {
                const uint32_t it = (uint32_t)itl;
                const long tile = mt0 + (itl / p.KC) * mstride;
                const int kc = (int)(itl % p.KC);
                const long m0 = tile * BM;
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                const int kbase = kc * BK + k4 * 4;
                const bool tr = TR_ON(blockIdx.x == 0 && t == 0 && grp == 0);
                const long long c0 = TR_CLOCK(tr);
                float sc[4], sh[4];
                if (p.a_scale) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = kbase + j < p.K;
                        sc[j] = ok ? __ldg(p.a_scale + kbase + j) : 0.f;
                        sh[j] = ok ? __ldg(p.a_shift + kbase + j) : 0.f;
                    }
                }
                float4 v[8];
                if (p.a_tma) {
                    const int rs = it % p.raw_slots;
                    const uint32_t rph = (it / p.raw_slots) & 1;
                    mbar_wait(&raw_full[rs], rph);
                    const unsigned char *slot = raw + (size_t)rs * RAW_BYTES;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        v[i] = *reinterpret_cast<const float4 *>(slot + (r0 + 16 * i) * 128 + k4 * 16);
                } else {
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
                }
                if (p.a_scale) {
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
                const long long c1 = TR_CLOCK(tr);
                mbar_wait(&empty[s], ph ^ 1);
                const long long c2 = TR_CLOCK(tr);
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
                // The raw slot goes back to the TMA loader only here: every value read from it has
                // been consumed (the stores above depend on them) and the proxy fence orders this
                // thread's generic-proxy reads of the slot before the async-proxy refill.  Releasing
                // it right after issuing the ld.shared let the refill overtake the reads (rare,
                // timing dependent: rows of the NEXT tile showed up in the accumulator).
                if (p.a_tma) mbar_arrive(&raw_empty[it % p.raw_slots]);
                if (tr) {
                    const long long c3 = TR_CLOCK(true);
                    TR_ADD(3, c1 - c0);
                    TR_ADD(4, c2 - c1);
                    TR_ADD(5, c3 - c2);
                }
            }
        }
    } else if (warp == W_ALOAD) {
        // ================================ A loader (TMA) ================================
        // one 2-D tensor load per 128 x 32 chunk into the raw ring, running raw_slots chunks ahead
        if (lane == 0 && p.a_tma) {
            uint32_t it = 0;
            for (long tile = mt0; tile < num_tiles; tile += mstride) {
                for (int kc = 0; kc < p.KC; ++kc, ++it) {
                    const int rs = it % p.raw_slots;
                    const uint32_t rph = (it / p.raw_slots) & 1;
                    mbar_wait(&raw_empty[rs], rph ^ 1);
                    mbar_expect_tx(&raw_full[rs], RAW_BYTES);
                    tma_load_2d(raw + (size_t)rs * RAW_BYTES, &tmA, kc * BK, (int)(tile * BM),
                                &raw_full[rs]);
                }
            }
        }
    } else if (warp == W_BLOAD) {
        // ================================ B loader ================================
        if (lane == 0 && p.b_res) {
            for (int kc = 0; kc < p.KC; ++kc) {
                mbar_expect_tx(&bfull[kc], 2 * b_bytes);
                bulk_g2s(bres + (size_t)kc * 2 * b_bytes,
                         image + (size_t)kc * 2 * b_bytes,
                         2 * b_bytes, &bfull[kc]);
            }
        } else if (lane == 0) {
            uint32_t it = 0;
            for (long tile = mt0; tile < num_tiles; tile += mstride) {
                for (int kc = 0; kc < p.KC; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    const long long c0 = TR_CLOCK(blockIdx.x == 0);
                    mbar_wait(&empty[s], ph ^ 1);
                    if (TR_ON(blockIdx.x == 0)) TR_ADD(8, TR_CLOCK(true) - c0);
                    unsigned char *b_hi = stage_base + (size_t)s * stage_bytes + 2 * a_bytes;
                    mbar_expect_tx(&full[s], 2 * b_bytes);
                    bulk_g2s(b_hi,
                             image + (size_t)kc * 2 * b_bytes,
                             2 * b_bytes, &full[s]);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================================ MMA issuer ================================
        {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc(p.Npad), idesc2 = make_idesc(2 * p.Npad);
            uint32_t it = 0, tcnt = 0;
            const bool tr = TR_ON(blockIdx.x == 0);
            const long long k0c = TR_CLOCK(tr);
            long long w_full = 0, w_issue = 0, w_acc = 0;
            const int kh = p.ksplit ? p.KC / 2 : p.KC;  // chunks per accumulator set
            for (long tile = mt0; tile < num_tiles; tile += mstride, ++tcnt) {
                // two accumulator sets: ping-pong over tiles, or (ksplit) the two K halves of one tile
                const uint32_t acc = p.ksplit ? 0 : (tcnt & 1), aph = p.ksplit ? (tcnt & 1) : ((tcnt >> 1) & 1);
                const long long ca = TR_CLOCK(tr);
                mbar_wait(&acc_empty[acc], aph ^ 1);
                if (tr) w_acc += TR_CLOCK(true) - ca;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kc = 0; kc < p.KC; ++kc, ++it) {
                    const uint32_t set = p.ksplit ? (kc >= kh ? 1u : 0u) : acc;
                    const int kcl = kc >= kh ? kc - kh : kc;
                    const uint32_t d = tmem_base + set * (uint32_t)(2 * Nacc);  // main
                    const uint32_t dc = d + (uint32_t)Nacc;                     // corrections
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    const long long cw = TR_CLOCK(tr);
                    mbar_wait(&full[s], ph);
                    if (p.b_res && tcnt == 0) mbar_wait(&bfull[kc], 0);
                    const long long ci = TR_CLOCK(tr);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(stage_base + (size_t)s * stage_bytes);
                    const uint32_t b_hi = p.b_res ? smem_u32(bres + (size_t)kc * 2 * b_bytes)
                                                  : a_hi + 2 * a_bytes;
                    const uint64_t dah = make_desc(a_hi), dal = make_desc(a_hi + a_bytes);
                    const uint64_t dbh = make_desc(b_hi), dbl = make_desc(b_hi + b_bytes);
                    if (leader) {
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2);  // 32 bytes per K=8 step, >>4
                        const uint32_t accum = (kcl > 0 || kk > 0) ? 1u : 0u;
                        if (p.stack) {
                            // the lo image follows the hi image row for row (Npad % 8 == 0), so [B_hi ; B_lo]
                            // is one K-major operand of 2*Npad rows: a_hi x [b_hi ; b_lo] fills main | corr
                            // (corr starts at column Nacc == Npad), a_lo x b_hi adds the second correction
                            umma_tf32(d, dah + adv, dbh + adv, idesc2, accum);
                            umma_tf32(dc, dal + adv, dbh + adv, idesc, 1u);
                        } else {
                            umma_tf32(d, dah + adv, dbh + adv, idesc, accum);
                            umma_tf32(dc, dal + adv, dbh + adv, idesc, accum);
                            umma_tf32(dc, dah + adv, dbl + adv, idesc, 1u);
                        }
                    }
                    umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
                    }
                    __syncwarp();
                    if (tr) {
                        w_full += ci - cw;
                        w_issue += TR_CLOCK(true) - ci;
                    }
                }
                if (leader) umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
                __syncwarp();
            }
            if (tr) {
                TR_ADD(0, w_full);
                TR_ADD(1, w_issue);
                TR_ADD(2, w_acc);
                TR_ADD(9, TR_CLOCK(true) - k0c);
                TR_ADD(10, it);
                TR_ADD(11, tcnt);
            }
        }
    } else if (warp < W_EPI) {
        // ================================ epilogue (warps 0-7) ================================
        // Warp w reads the TMEM lane quadrant w & 3 (rows) and the 32-column blocks cb with
        // cb & 1 == w >> 2.  tcgen05.ld (lane = row) -> accumulators + bias -> 128B-swizzled staging
        // tile (32 rows x 32 columns) -> one TMA tensor store per block (or, without a Y tensor map,
        // coalesced st.global of the columns read back from the tile); the BatchNorm column
        // statistics read the same staging tile column-wise.
        const int q4 = warp & 3, wg = warp >> 2;
        unsigned char *buf = epi_b + warp * 4096;
        const int nblk = (Nv + 31) / 32;
        // the bias vector is the epilogue's business alone: loaded here, behind a barrier of the epilogue warps only,
        // so that the loader / transform / MMA roles do not sit out a global-load round trip at kernel start
        if (threadIdx.x < 128)
            sbias[threadIdx.x] = (p.bias && (int)threadIdx.x < Nv) ? __ldg(p.bias + n0 + threadIdx.x) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * W_EPI) : "memory");
        // BatchNorm statistics: per lane-column fp64 sums of (y - c) and (y - c)^2 with a constant
        // shift c (the first value this lane sees in the column) so that neither the fp32 partial
        // sums over 32 rows nor the final variance suffer cancellation; un-shifted once at the end.
        double ssum[2], ssq[2];
        float cshift[2];
        long nrows = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ssum[i] = ssq[i] = 0.0;
            cshift[i] = 0.f;
        }
        // 16 accumulator columns of this lane's row: main + corrections (round-to-nearest), summed
        // over both K halves when the contraction was split
        auto load_sum16 = [&](uint32_t taddr, float (&o)[16]) {
            uint32_t r[16], rc[16];
            tmem_ld16_nowait(taddr, r);
            tmem_ld16_nowait(taddr + (uint32_t)Nacc, rc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
            if (p.ksplit) {
                tmem_ld16_nowait(taddr + (uint32_t)(2 * Nacc), r);
                tmem_ld16_nowait(taddr + (uint32_t)(3 * Nacc), rc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] += __uint_as_float(r[j]) + __uint_as_float(rc[j]);
            }
        };
        uint32_t tcnt = 0;
        for (long tile = mt0; tile < num_tiles; tile += mstride, ++tcnt) {
            // narrow outputs: a tile's epilogue is one ~2 k-cycle dependent chain (accumulator wait, tcgen05.ld,
            // staging, tensor store, statistics); two warp groups on alternate tiles keep two in flight
            if (p.epi_alt && (int)(tcnt & 1) != wg) continue;
            const uint32_t acc = p.ksplit ? 0 : (tcnt & 1), aph = p.ksplit ? (tcnt & 1) : ((tcnt >> 1) & 1);
            const long m0 = tile * BM + q4 * 32;
            const bool tr = TR_ON(blockIdx.x == 0 && threadIdx.x == 0);
            const long long ce0 = TR_CLOCK(tr);
            mbar_wait(&acc_full[acc], aph);
            const long long ce1 = TR_CLOCK(tr);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long rows_left = p.M - m0;
            const int nv = rows_left >= 32 ? 32 : (int)rows_left;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int cb = p.epi_alt ? ci : 2 * ci + wg;
                if (cb >= nblk) break;
                const uint32_t ta = tmem_base + ((uint32_t)(q4 * 32) << 16) +
                                    acc * (uint32_t)(2 * Nacc) + cb * 32;
                const int col = cb * 32 + lane;
                const bool col_ok = col < Nv;
                const bool do_stats = p.stats_sum && col_ok && rows_left > 0;
                // the tensor store issued from the staging tile one block ago has read it
                if (p.y_tma && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float a16[16];
                    load_sum16(ta + half * 16, a16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b4 =
                            *reinterpret_cast<const float4 *>(sbias + cb * 32 + half * 16 + q * 4);
                        float4 o;  // accumulators + bias
                        o.x = a16[q * 4] + b4.x;
                        o.y = a16[q * 4 + 1] + b4.y;
                        o.z = a16[q * 4 + 2] + b4.z;
                        o.w = a16[q * 4 + 3] + b4.w;
                        const int c16 = half * 4 + q;
                        *reinterpret_cast<float4 *>(buf + lane * 128 + ((c16 ^ (lane & 7)) << 4)) = o;
                    }
                }
                if (p.y_tma) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && rows_left > 0) {
                        tma_store_2d(&tmY, n0 + cb * 32, (int)m0, buf);  // rows >= M, cols >= N clipped
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                    __syncwarp();
                }
                if (do_stats || !p.y_tma) {
                    // column `lane` of the tile, rows 0..31 (conflict-free: a row is one 128-byte line)
                    const unsigned char *colp = buf + (lane & 3) * 4;
                    const int c16 = lane >> 2;
                    float vals[32];
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr)
                        vals[rr] = *reinterpret_cast<const float *>(colp + rr * 128 + (((c16 ^ (rr & 7)) & 7) << 4));
                    if (!p.y_tma && col_ok && rows_left > 0) {
                        float *yp = p.Y + m0 * p.ldy + n0 + col;
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr)
                            if (rr < nv) yp[(long)rr * p.ldy] = vals[rr];
                    }
                    if (do_stats) {
                        if (nrows == 0) cshift[ci] = vals[0];  // this warp's first tile
                        const float c0 = cshift[ci];
                        float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) {
                            const float dv = rr < nv ? vals[rr] - c0 : 0.f;
                            p1[rr & 3] += dv;
                            p2[rr & 3] = __fmaf_rn(dv, dv, p2[rr & 3]);
                        }
                        ssum[ci] += (double)((p1[0] + p1[1]) + (p1[2] + p1[3]));
                        ssq[ci] += (double)((p2[0] + p2[1]) + (p2[2] + p2[3]));
                    }
                    if (!p.y_tma) __syncwarp();  // all column reads done before the tile is rewritten
                }
            }
            nrows += rows_left >= 32 ? 32 : (rows_left > 0 ? rows_left : 0);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[acc]);
            if (tr) {
                TR_ADD(6, ce1 - ce0);
                TR_ADD(7, TR_CLOCK(true) - ce1);
            }
        }
        if (p.y_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (p.stats_sum) {
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int cb = p.epi_alt ? ci : 2 * ci + wg;
                const int col = cb * 32 + lane;
                if (cb < nblk && col < Nv && nrows > 0) {
                    const double c = (double)cshift[ci], n = (double)nrows;
                    atomicAdd(p.stats_sum + n0 + col, ssum[ci] + n * c);
                    atomicAdd(p.stats_sq + n0 + col, ssq[ci] + 2.0 * c * ssum[ci] + n * c * c);
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    pdl_trigger_late();  // all tiles of this CTA are done: what follows is the TMEM release and the statistics finalize
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(ncols)
                     : "memory");
    }
    if (p.fin_scale) {
        // every CTA has issued its fp64 statistics atomics (the barrier above); the last one to get here
        // sees all of them (fence + counter: the classic last-block pattern) and finalises the layer
        // (the flag lives in the spare word next to the TMEM address: a static __shared__ variable would push
        //  static + dynamic shared memory past the 227 KB this kernel opts in to)
        volatile uint32_t *s_last = tmem_slot + 1;
        if (threadIdx.x == 0) {
            __threadfence();
            *s_last = atomicAdd(p.fin_done, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (*s_last) {
            __threadfence();
            const double Md = (double)p.M;
            for (int c = threadIdx.x; c < p.Ntot; c += THREADS) {
                const double mean = __ldcg(p.stats_sum + c) / Md;
                double var = __ldcg(p.stats_sq + c) / Md - mean * mean;
                if (var < 0.0) var = 0.0;
                const double rstd = 1.0 / sqrt(var + (double)p.fin_eps);
                const double g = (double)__ldg(p.fin_gamma + c);
                p.fin_scale[c] = (float)(g * rstd);
                p.fin_shift[c] = (float)((double)__ldg(p.fin_beta + c) - mean * g * rstd);
                p.fin_saved[c] = (float)mean;
                p.fin_saved[p.Ntot + c] = (float)rstd;
                if (p.fin_mm) {
                    const double uv = p.fin_unbiased && p.M > 1 ? var * (Md / (Md - 1.0)) : var;
                    p.fin_mm[c] = (float)((double)p.fin_mm[c] - ((double)p.fin_mm[c] - mean) * (1.0 - (double)p.fin_decay));
                    p.fin_mv[c] = (float)((double)p.fin_mv[c] - ((double)p.fin_mv[c] - uv) * (1.0 - (double)p.fin_decay));
                }
            }
        }
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

// cuTensorMapEncodeTiled is reached through the runtime's driver entry point query: the library
// does not link libcuda (it must load on machines without a driver for the ABI tests).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// row-major fp32 matrix [rows x cols], leading dimension ld (floats): box = box_rows x 32 columns
static bool make_tensor_map(CUtensorMap *tm, const float *base, long rows, int cols, int ld,
                            int box_rows, bool swizzle128) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (ld % 4) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || rows <= 0 || cols <= 0)
        return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1u, 1u};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t image_bytes(int K, int N) {
    const int KC = (K + BK - 1) / BK;
    if (N > 128) return (size_t)((N + 127) / 128) * KC * 2 * 128 * 128;
    const int Npad = (N + 15) & ~15;
    return (size_t)KC * 2 * Npad * 128;
}

// Y[M, 0:Ntot]: nchunks column blocks of Nc columns (Nc == 128 when nchunks > 1, the last one may be
// narrower; a single block has Nc = Ntot <= 128)
static int run_chunk(long M, int K, int Nc, int nchunks, int Ntot, const float *A, int lda, const float *a_scale,
                     const float *a_shift, int a_relu, const float *bsrc, long s_n, long s_k,
                     const float *bias, float *Y, int ldy, double *stats_sum, double *stats_sq,
                     float *ws, bool image_ready, const pn2_bn_finalize *fin, cudaStream_t st) {
    Params p;
    p.fin_gamma = p.fin_beta = nullptr;
    p.fin_mm = p.fin_mv = p.fin_scale = p.fin_shift = p.fin_saved = nullptr;
    p.fin_done = nullptr;
    p.fin_eps = p.fin_decay = 0.f;
    p.fin_unbiased = 0;
    if (fin && stats_sum) {
        p.fin_gamma = fin->gamma;
        p.fin_beta = fin->beta;
        p.fin_mm = fin->moving_mean;
        p.fin_mv = fin->moving_var;
        p.fin_scale = fin->scale;
        p.fin_shift = fin->shift;
        p.fin_saved = fin->saved;
        p.fin_done = fin->counter;
        p.fin_eps = fin->eps;
        p.fin_decay = fin->decay;
        p.fin_unbiased = fin->unbiased_moving;
    }
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
    if (p.KC > MAX_KC || Nc > 128 || nchunks < 1 || (nchunks > 1 && Nc != 128)) return PN2_EUNSUPPORTED;
    p.nchunks = nchunks;
    p.Ntot = Ntot;
    p.ksplit = p.KC > 16 ? 1 : 0;  // K > 512: two accumulator sets bound the truncating accumulation
    static const bool alt_ok = !(getenv("PN2_TC_EPI_ALT") && getenv("PN2_TC_EPI_ALT")[0] == '0');
    p.epi_alt = (alt_ok && !p.ksplit && nchunks == 1 && Nc <= 32) ? 1 : 0;
    // PN2_TC_STACK=0: three separate MMAs per K=8 step (A/B switch)
    static const bool stack_ok = !(getenv("PN2_TC_STACK") && getenv("PN2_TC_STACK")[0] == '0');
    p.stack = (stack_ok && (p.Npad % 32) == 0 && 2 * p.Npad <= 256) ? 1 : 0;

    // PN2_TC_TMA bit mask (diagnostics): 1 = tensor loads for A, 2 = tensor stores for Y; default 3
    static const int tma_mask = getenv("PN2_TC_TMA") ? atoi(getenv("PN2_TC_TMA")) : 3;
    CUtensorMap tmA, tmY;
    memset(&tmA, 0, sizeof(tmA));
    memset(&tmY, 0, sizeof(tmY));
    p.a_tma = ((tma_mask & 1) && M < (1l << 31) && make_tensor_map(&tmA, A, M, K, lda, BM, false)) ? 1 : 0;
    p.y_tma = ((tma_mask & 2) && M < (1l << 31) && make_tensor_map(&tmY, Y, M, Ntot, ldy, 32, true)) ? 1 : 0;

    // shared memory plan (227 KB usable per CTA): raw ring (4 slots if possible), >= 2 MMA stages,
    // weights resident when they fit next to that
    const size_t a_stage = 2 * (size_t)BM * 128, b_chunk = 2 * (size_t)p.Npad * 128;
    const size_t fixed = EPI_BYTES + 512 + (2 * MAX_STAGES + 4 + MAX_KC + 2 * MAX_RAW) * 8 + 16;
    const size_t budget = 227 * 1024 - fixed;
    static const int force_stream = getenv("PN2_TC_STREAM_B") ? atoi(getenv("PN2_TC_STREAM_B")) : 0;
    const size_t bres = (size_t)p.KC * b_chunk;
    // (an even number of raw slots: the two transform groups must each own fixed slots, otherwise
    //  a group could test a slot's mbarrier one phase early)
    int raw_slots = p.a_tma ? MAX_RAW : 0, stages = 0;
    p.b_res = 0;
    for (;; raw_slots -= 2) {
        const size_t rawb = (size_t)raw_slots * RAW_BYTES;
        if (!force_stream && bres + 2 * a_stage + rawb <= budget) {
            p.b_res = 1;
            stages = (int)((budget - bres - rawb) / a_stage);
            break;
        }
        if (2 * (a_stage + b_chunk) + rawb <= budget) {
            stages = (int)((budget - rawb) / (a_stage + b_chunk));
            break;
        }
        if (raw_slots <= (p.a_tma ? 2 : 0)) {  // last resort: a single stage
            stages = (int)((budget - rawb) / (a_stage + b_chunk));
            break;
        }
    }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 1) return PN2_EUNSUPPORTED;
    p.stages = stages;
    p.raw_slots = raw_slots;
    const size_t stage_bytes = p.b_res ? a_stage : a_stage + b_chunk;
    const size_t smem = (p.b_res ? bres : 0) + (size_t)stages * stage_bytes +
                        (size_t)raw_slots * RAW_BYTES + fixed;

    int rc = PN2_OK;
    if (!image_ready) {  // per-call image (callers that did not batch the preparation)
        const long total = (long)nchunks * p.KC * p.Npad * BK;
        int pb = (int)((total + 255) / 256);
        if (pb > 148 * 8) pb = 148 * 8;
        launch_k(tc_prep_b_kernel, pb, 256, 0, st, Nc, Ntot, K, p.Npad, p.KC, nchunks, bsrc, s_n, s_k, ws);
        rc = finish_launch();
        if (rc) return rc;
    }

    // opt in to the full 227 KB once per device (not per launch: no driver call on the hot path,
    // nothing that could disturb a stream capture)
    rc = opt_in_smem(reinterpret_cast<const void *>(tc_gemm_kernel), 0);
    if (rc) return rc;
    const long tiles = (M + BM - 1) / BM;
    long per_chunk = num_sms() / nchunks;  // persistent CTAs per column block
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > tiles) per_chunk = tiles;
    const int grid = (int)per_chunk * nchunks;
    launch_k(tc_gemm_kernel, grid, THREADS, smem, st, tmA, tmY, p);
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
// transform warps copy rows (BN affine + ReLU, hi/lo split) without any transpose.
//
// Work decomposition: one launch, persistent CTAs over units (row segment, 128-feature block,
// 128-column block).  The row segments are sized so that the units fill the machine (short
// problems get 128-row segments, long ones up to 512 rows); each unit accumulates in its own TMEM
// accumulator pair (double buffered) and is flushed to dW with vectorised fp32 reductions
// (red.global.add.v4.f32) straight from the TMEM registers.  Segments of at most 512 rows bound
// the number of truncating tensor-core accumulations (64 per accumulator).
// Data path: TMA tensor loads of raw X / dY row blocks into a 2-slot ring (the loads run ahead
// of the math, nothing is staged through registers) -> two transform groups of 4 warps -> 2 MMA
// stages.  A stage holds 32 / 64 / 128 contraction rows depending on the operand widths so that
// every transform thread always moves 8 + 8 float4.
// =====================================================================================================
namespace tcw {
using namespace tc;

constexpr int W_NPG = 2;                        // transform warpgroups: ONE group of 8 warps works on a stage (two groups
                                                // of 4 on alternating stages measured slower: 1.08 vs 0.955 ms per step --
                                                // this transform is bound by shared-memory throughput, not by latency)
constexpr int W_LOAD = 4 + 4 * W_NPG;           // TMA loader warp
constexpr int W_WMMA = W_LOAD + 1;              // MMA / TMEM warp
constexpr int W_THREADS = 32 * (W_WMMA + 1);    // warps 0-3 epilogue, 4-11 transform, loader, MMA
constexpr int W_MAXSEG = 512;   // contraction rows per accumulator segment (accuracy bound)
constexpr int W_FEAT = 128;     // features (rows of dW) per unit = UMMA M
constexpr int W_RAW = 3;        // raw ring slots (all 256 transform threads walk the stages together,
                                // so every thread tests every slot's mbarrier in order)
constexpr int W_TT = 128 * W_NPG;               // transform threads
constexpr int W_STAGES = 2;     // MMA stages

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
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

struct WParams {
    long M, nseg, units;
    int K, N, Kdo, ldw, a_relu;
    int MG, NG;      // 32-feature groups of X / 32-column groups of dY held per stage (1..4)
    int rows;        // contraction rows per stage: 128 / max(MG,NG) rounded to 32, 64, 128
    int seg_rows;    // rows per unit (multiple of rows, <= W_MAXSEG)
    int KB, NB;      // feature blocks / column blocks
    int stack;       // dY stage laid out as [hi | lo] MN-groups of one 2*NG-group operand: 2 MMAs per step
    const float *a_scale, *a_shift;
    float *dW;
};

__global__ void __launch_bounds__(W_THREADS, 1)
    tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                    const WParams p) {
    pdl_trigger();
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    }
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve: W_STAGES x [X_hi | X_lo | G_hi | G_lo], W_RAW x [X raw | G raw], barriers
    const uint32_t a_bytes = (uint32_t)p.rows * p.MG * 128;
    const uint32_t b_bytes = (uint32_t)p.rows * p.NG * 128;
    const uint32_t raw_bytes = a_bytes + b_bytes;
    const uint32_t stage_bytes = 2 * raw_bytes;
    unsigned char *raw = smem + (size_t)W_STAGES * stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(raw + (size_t)W_RAW * raw_bytes);
    uint64_t *full = bars, *empty = bars + W_STAGES;           // MMA stages
    uint64_t *raw_full = bars + 2 * W_STAGES, *raw_empty = raw_full + W_RAW;
    uint64_t *acc_full = raw_empty + W_RAW, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Nacc = 32 * p.NG;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(4 * Nacc)) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < W_STAGES; ++s) {
            mbar_init(&full[s], W_TT);
            mbar_init(&empty[s], 1);
        }
        for (int r = 0; r < W_RAW; ++r) {
            mbar_init(&raw_full[r], 1);
            mbar_init(&raw_empty[r], W_TT);
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
    pdl_wait();  // everything above (barrier init, TMEM allocation) overlaps the previous kernel's tail
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int upb = p.KB * p.NB;  // units per row segment; consecutive units share their rows (L2)
    // stages of unit u: rows [seg*seg_rows, min(M, (seg+1)*seg_rows)) in steps of p.rows
    auto unit_stages = [&](long u) -> int {
        const long seg0 = (u / upb) * (long)p.seg_rows;
        const long len = min((long)p.seg_rows, p.M - seg0);
        return (int)((len + p.rows - 1) / p.rows);
    };

    if (warp >= 4 && warp < W_LOAD) {
        // ================================ transform ================================
        const int tt = threadIdx.x - 128;                  // 0 .. W_TT-1
        const int aw4 = 8 * p.MG, bw4 = 8 * p.NG;         // float4 per raw row
        const int a_iters = (p.rows * p.MG * 8) / W_TT;    // <= 4
        const int b_iters = (p.rows * p.NG * 8) / W_TT;    // <= 4
        const int ac4 = tt % aw4, ar0 = tt / aw4, ar_step = W_TT / aw4;  // W_TT % aw4 == 0
        uint32_t it = 0;
        for (long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const long seg0 = (u / upb) * (long)p.seg_rows;
            const int k0 = (int)((u % upb) / p.NB) * W_FEAT;
            const int nst = unit_stages(u);
            const int kf = k0 + 4 * ac4;
            float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
            if (p.a_scale) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (kf + j < p.Kdo) {
                        sc[j] = __ldg(p.a_scale + kf + j);
                        sh[j] = __ldg(p.a_shift + kf + j);
                    }
            }
            for (int sidx = 0; sidx < nst; ++sidx, ++it) {
                const int s = it % W_STAGES;
                const uint32_t ph = (it / W_STAGES) & 1;
                const int rs = it % W_RAW;
                const uint32_t rph = (it / W_RAW) & 1;
                const long mbase = seg0 + (long)sidx * p.rows;
                const unsigned char *rx = raw + (size_t)rs * raw_bytes;
                const unsigned char *rg = rx + a_bytes;
                float4 va[4], vb[4];
                mbar_wait(&raw_full[rs], rph);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < a_iters)
                        va[i] = *reinterpret_cast<const float4 *>(rx + (size_t)(tt + W_TT * i) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q < b_iters)
                        vb[q] = *reinterpret_cast<const float4 *>(rg + (size_t)(tt + W_TT * q) * 16);
                if (p.a_scale) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (i < a_iters) {
                            const bool row_ok = (mbase + ar0 + ar_step * i) < p.M;
                            float *e = reinterpret_cast<float *>(&va[i]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float y = __fmaf_rn(e[j], sc[j], sh[j]);
                                if (p.a_relu) y = fmaxf(y, 0.f);
                                e[j] = (row_ok && kf + j < p.Kdo) ? y : 0.f;
                            }
                        }
                    }
                }
                unsigned char *st_base = smem + (size_t)s * stage_bytes;
                mbar_wait(&empty[s], ph ^ 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < a_iters) {
                        const uint32_t off = mn_offset(ar0 + ar_step * i, ac4, p.MG);
                        float4 hi, lo;
                        hi.x = tf32_rna(va[i].x); lo.x = va[i].x - hi.x;
                        hi.y = tf32_rna(va[i].y); lo.y = va[i].y - hi.y;
                        hi.z = tf32_rna(va[i].z); lo.z = va[i].z - hi.z;
                        hi.w = tf32_rna(va[i].w); lo.w = va[i].w - hi.w;
                        *reinterpret_cast<float4 *>(st_base + off) = hi;
                        *reinterpret_cast<float4 *>(st_base + a_bytes + off) = lo;
                    }
                }
                unsigned char *b_hi = st_base + 2 * a_bytes;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < b_iters) {
                        const int e = tt + W_TT * q;
                        float4 hi, lo;
                        hi.x = tf32_rna(vb[q].x); lo.x = vb[q].x - hi.x;
                        hi.y = tf32_rna(vb[q].y); lo.y = vb[q].y - hi.y;
                        hi.z = tf32_rna(vb[q].z); lo.z = vb[q].z - hi.z;
                        hi.w = tf32_rna(vb[q].w); lo.w = vb[q].w - hi.w;
                        if (p.stack) {  // one operand of 2*NG MN-groups per k-group: [hi groups | lo groups]
                            *reinterpret_cast<float4 *>(b_hi + mn_offset(e / bw4, e % bw4, 2 * p.NG)) = hi;
                            *reinterpret_cast<float4 *>(b_hi + mn_offset(e / bw4, e % bw4 + bw4, 2 * p.NG)) = lo;
                        } else {
                            const uint32_t off = mn_offset(e / bw4, e % bw4, p.NG);
                            *reinterpret_cast<float4 *>(b_hi + off) = hi;
                            *reinterpret_cast<float4 *>(b_hi + b_bytes + off) = lo;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&full[s]);
                mbar_arrive(&raw_empty[rs]);  // after the fence: see the forward kernel
            }
        }
    } else if (warp == W_LOAD) {
        // ================================ TMA loader ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (long u = blockIdx.x; u < p.units; u += gridDim.x) {
                const long seg0 = (u / upb) * (long)p.seg_rows;
                const int k0 = (int)((u % upb) / p.NB) * W_FEAT;
                const int n0 = (int)(u % p.NB) * 128;
                const int nst = unit_stages(u);
                for (int sidx = 0; sidx < nst; ++sidx, ++it) {
                    const int rs = it % W_RAW;
                    const uint32_t rph = (it / W_RAW) & 1;
                    const int m = (int)(seg0 + (long)sidx * p.rows);
                    mbar_wait(&raw_empty[rs], rph ^ 1);
                    mbar_expect_tx(&raw_full[rs], raw_bytes);
                    unsigned char *dst = raw + (size_t)rs * raw_bytes;
                    tma_load_2d(dst, &tmX, k0, m, &raw_full[rs]);          // rows/features out of range: 0
                    tma_load_2d(dst + a_bytes, &tmG, n0, m, &raw_full[rs]);
                }
            }
        }
    } else if (warp == W_WMMA) {
        // ================================ MMA issuer ================================
        {
            const bool leader = elect_one();
            // UMMA M is always 128: feature groups >= MG read shared memory of the neighbouring
            // atoms and produce accumulator rows that nobody reads
            const uint32_t idesc = make_idesc(Nacc) | (1u << 15) | (1u << 16);  // A, B MN-major
            const uint32_t idesc2 = make_idesc(2 * Nacc) | (1u << 15) | (1u << 16);
            uint32_t it = 0, ucnt = 0;
            const bool tr = TR_ON(blockIdx.x == 0);
            const long long k0c = TR_CLOCK(tr);
            long long w_full = 0, w_issue = 0, w_acc = 0;
            const int ksteps = p.rows / 8;
            for (long u = blockIdx.x; u < p.units; u += gridDim.x, ++ucnt) {
                const int nst = unit_stages(u);
                const uint32_t acc = ucnt & 1, aph = (ucnt >> 1) & 1;
                const long long ca = TR_CLOCK(tr);
                mbar_wait(&acc_empty[acc], aph ^ 1);
                if (tr) w_acc += TR_CLOCK(true) - ca;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * (uint32_t)(2 * Nacc);
                const uint32_t dc = d + (uint32_t)Nacc;
                for (int sidx = 0; sidx < nst; ++sidx, ++it) {
                    const int s = it % W_STAGES;
                    const uint32_t ph = (it / W_STAGES) & 1;
                    const long long cw = TR_CLOCK(tr);
                    mbar_wait(&full[s], ph);
                    const long long ci = TR_CLOCK(tr);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t b_hi = a_hi + 2 * a_bytes;
                    if (leader) {
                    for (int ks = 0; ks < ksteps; ++ks) {
                        // one K=8 MMA spans two 4-row k-groups: LBO = next MN group (512 B),
                        // SBO = next k-group (groups * 512 B)
                        const uint32_t ao = ks * p.MG * 1024;
                        const uint64_t dah = make_desc_mn(a_hi + ao, 512, p.MG * 512);
                        const uint64_t dal = make_desc_mn(a_hi + a_bytes + ao, 512, p.MG * 512);
                        const uint32_t accum = (sidx > 0 || ks > 0) ? 1u : 0u;
                        if (p.stack) {
                            // x_hi^T [g_hi | g_lo] fills main | corr in one MMA of 2*Nacc columns, x_lo^T g_hi
                            // (the first NG MN-groups of the same operand) adds the second correction
                            const uint64_t dbs = make_desc_mn(b_hi + ks * p.NG * 2048, 512, p.NG * 1024);
                            umma_tf32(d, dah, dbs, idesc2, accum);
                            umma_tf32(dc, dal, dbs, idesc, 1u);
                        } else {
                            const uint32_t bo = ks * p.NG * 1024;
                            const uint64_t dbh = make_desc_mn(b_hi + bo, 512, p.NG * 512);
                            const uint64_t dbl = make_desc_mn(b_hi + b_bytes + bo, 512, p.NG * 512);
                            umma_tf32(d, dah, dbh, idesc, accum);
                            umma_tf32(dc, dal, dbh, idesc, accum);
                            umma_tf32(dc, dah, dbl, idesc, 1u);
                        }
                    }
                    umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    if (tr) {
                        w_full += ci - cw;
                        w_issue += TR_CLOCK(true) - ci;
                    }
                }
                if (leader) umma_commit(&acc_full[acc]);
                __syncwarp();
            }
            if (tr) {
                TR_ADD(0, w_full);
                TR_ADD(1, w_issue);
                TR_ADD(2, w_acc);
                TR_ADD(9, TR_CLOCK(true) - k0c);
                TR_ADD(10, it);
                TR_ADD(11, ucnt);
            }
        }
    } else if (warp < 4) {
        // ================================ epilogue ================================
        // lane = feature row of dW; 16 consecutive columns per tcgen05.ld -> 4 vector reductions
        uint32_t ucnt = 0;
        for (long u = blockIdx.x; u < p.units; u += gridDim.x, ++ucnt) {
            const int k0 = (int)((u % upb) / p.NB) * W_FEAT;
            const int n0 = (int)(u % p.NB) * 128;
            const uint32_t acc = ucnt & 1, aph = (ucnt >> 1) & 1;
            mbar_wait(&acc_full[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int k = k0 + warp * 32 + lane;
            if (warp < p.MG) {
                float *wrow = p.dW + (long)k * p.ldw;
                for (int cb = 0; cb < p.NG; ++cb) {
                    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                        acc * (uint32_t)(2 * Nacc) + cb * 32;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t r[16], rc[16];
                        tmem_ld16_nowait(ta + half * 16, r);
                        tmem_ld16_nowait(ta + (uint32_t)Nacc + half * 16, rc);
                        tmem_ld_wait();
                        if (k < p.Kdo) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int n = n0 + cb * 32 + half * 16 + q * 4;
                                const float o0 = __uint_as_float(r[q * 4]) + __uint_as_float(rc[q * 4]);
                                const float o1 = __uint_as_float(r[q * 4 + 1]) + __uint_as_float(rc[q * 4 + 1]);
                                const float o2 = __uint_as_float(r[q * 4 + 2]) + __uint_as_float(rc[q * 4 + 2]);
                                const float o3 = __uint_as_float(r[q * 4 + 3]) + __uint_as_float(rc[q * 4 + 3]);
                                if (n + 3 < p.N) {
                                    red_add_v4(wrow + n, o0, o1, o2, o3);  // N % 4 == 0: 16-byte aligned
                                } else {
                                    if (n < p.N) atomicAdd(wrow + n, o0);
                                    if (n + 1 < p.N) atomicAdd(wrow + n + 1, o1);
                                    if (n + 2 < p.N) atomicAdd(wrow + n + 2, o2);
                                }
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[acc]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    pdl_trigger_late();
    if (warp == W_WMMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(ncols)
                     : "memory");
    }
}

// row-major fp32 matrix [rows x cols] (leading dimension ld): box = box_rows x box_cols, no swizzle
static bool make_map(CUtensorMap *tm, const float *base, long rows, int cols, int ld, int box_rows,
                     int box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (ld % 4) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || rows <= 0 || cols <= 0)
        return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1u, 1u};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int run(long M, int K, int Kdo, int N, const float *A, int lda, const float *a_scale,
               const float *a_shift, int a_relu, const float *dY, float *dW, cudaStream_t st) {
    WParams p;
    p.M = M;
    p.K = K;
    p.N = N;
    p.Kdo = Kdo;
    p.ldw = N;
    p.a_relu = a_relu;
    p.a_scale = a_scale;
    p.a_shift = a_shift;
    p.dW = dW;
    p.KB = (Kdo + W_FEAT - 1) / W_FEAT;
    p.NB = (N + 127) / 128;
    p.MG = Kdo >= W_FEAT ? 4 : (Kdo + 31) / 32;
    if (p.MG == 3) p.MG = 4;  // transform threads keep fixed feature columns: 128 % (8*MG) == 0
    p.NG = N >= 128 ? 4 : (N + 31) / 32;
    const int mx = p.MG > p.NG ? p.MG : p.NG;
    p.rows = mx == 1 ? 128 : (mx == 2 ? 64 : 32);
    // row segments: as long as allowed (<= 512 rows) while the units still fill the machine twice
    const int upb = p.KB * p.NB;
    const int min_seg = p.rows > 128 ? p.rows : 128;
    int seg = W_MAXSEG;
    for (int r = 2; r <= 64; r += 2) {
        long S = (long)num_sms() * r / upb;
        if (S < 1) S = 1;
        long sr = ((M + S - 1) / S + p.rows - 1) / p.rows * p.rows;
        if (sr < min_seg) sr = min_seg;
        if (sr <= W_MAXSEG) {
            seg = (int)sr;
            break;
        }
    }
    static const bool stack_ok = !(getenv("PN2_TC_STACK") && getenv("PN2_TC_STACK")[0] == '0');
    p.stack = stack_ok ? 1 : 0;  // 2 * Nacc <= 256 always (NG <= 4)
    p.seg_rows = seg;
    p.nseg = (M + seg - 1) / seg;
    p.units = p.nseg * upb;

    CUtensorMap tmX, tmG;
    if (!make_map(&tmX, A, M, Kdo, lda, p.rows, 32 * p.MG) ||
        !make_map(&tmG, dY, M, N, N, p.rows, 32 * p.NG))
        return PN2_EUNSUPPORTED;

    const size_t raw_bytes = (size_t)p.rows * (p.MG + p.NG) * 128;
    const size_t smem = (size_t)(2 * W_STAGES + W_RAW) * raw_bytes + (2 * W_STAGES + 2 * W_RAW + 4) * 8 + 16;
    int rc = tc::opt_in_smem(reinterpret_cast<const void *>(tc_wgrad_kernel), 1);
    if (rc) return rc;
    const int grid = (int)(p.units < num_sms() ? p.units : num_sms());
    launch_k(tc_wgrad_kernel, grid, W_THREADS, smem, st, tmX, tmG, p);
    return finish_launch();
}

}  // namespace tcw

int tc_linear_wgrad(long M, int K, int N, const float *A, int lda, const float *a_scale,
                    const float *a_shift, int a_relu, const float *dY, float *dW, bool force,
                    int *k_done, cudaStream_t st) {
    *k_done = 0;
    // tensor maps need 16-byte aligned rows (lda % 4, N % 4) and 16-byte aligned bases
    if (M < 512 || N < 16 || K < 4 || (N % 4) != 0 || (lda % 4) != 0 || M >= (1l << 31))
        return PN2_EUNSUPPORTED;
    // a narrow tail of features (K = 131 -> 3) would cost a full extra pass over dY on the
    // 128-feature MMA tile; it is left to the caller's fp32 kernel (k_done tells where it starts)
    int Kdo = K;
    static const bool tail_simt = getenv("PN2_WGRAD_TAIL") && getenv("PN2_WGRAD_TAIL")[0] == '1';
    if (tail_simt && !force && K > tcw::W_FEAT && (K % tcw::W_FEAT) < 32) Kdo = K - (K % tcw::W_FEAT);
    int rc = tcw::run(M, K, Kdo, N, A, lda, a_scale, a_shift, a_relu, dY, dW, st);
    if (rc == PN2_OK) *k_done = Kdo;
    return rc;
}

// Shapes worth the tensor cores: at least one full tile of rows, K and N not tiny.
// Each accumulator set sees at most 512 contraction terms (K > 512 is split in two halves), which
// keeps the truncating tensor-core accumulation inside the 1e-5 parity bar.
// Narrow layers run too (the 6-channel input layer, the 9-class head): a partial K chunk and the
// columns past N are zero-filled in shared memory.
static bool tc_shape_ok(long M, int K, int N) { return M >= 128 && K >= 1 && K <= 1024 && N >= 1; }

constexpr int TC_NCHUNK = 128;
static size_t tc_image_bytes(int K, int N) { return tc::image_bytes(K, N); }

// one buffer serves both orientations of a layer: forward (K x N) and dgrad (N x K)
size_t tc_workspace_bytes(int K, int N) {
    const size_t a = tc_image_bytes(K, N), b = tc_image_bytes(N, K);
    return a > b ? a : b;
}

// table entry of one image: forward orientation Bt(n,k) = W[k*N + n]; dgrad orientation (output
// columns = K, contraction over N) Bt(k,n) = W[k*N + n]
int tc_describe_image(int K, int N, bool dgrad, const float *W, float *image, pn2_linear_image *out) {
    const int Kc = dgrad ? N : K, Nout = dgrad ? K : N;  // contraction length, output columns
    if (Kc < 1 || Kc > 1024 || Nout < 1) return PN2_EUNSUPPORTED;
    const int nch = (Nout + TC_NCHUNK - 1) / TC_NCHUNK;
    const int Nc = nch == 1 ? Nout : TC_NCHUNK;
    out->src = W;
    out->image = image;
    out->s_n = dgrad ? N : 1;
    out->s_k = dgrad ? 1 : N;
    out->N = Nc;
    out->Ntot = Nout;
    out->K = Kc;
    out->Npad = (Nc + 15) & ~15;
    out->KC = (Kc + tc::BK - 1) / tc::BK;
    out->nchunks = nch;
    out->total = (long)nch * out->KC * out->Npad * tc::BK;
    return PN2_OK;
}

size_t tc_image_bytes_for(int K, int N, bool dgrad) {
    const int Kc = dgrad ? N : K, Nout = dgrad ? K : N;
    if (Kc < 1 || Kc > 1024 || Nout < 1) return 0;
    return tc_image_bytes(Kc, Nout);
}

int tc_prepare_images(int count, const pn2_linear_image *table_dev, cudaStream_t st) {
    if (count <= 0) return PN2_OK;
    dim3 grid(32, (unsigned)count);
    launch_k(tc::tc_prep_images_kernel, grid, 256, 0, st, table_dev);
    return finish_launch();
}

int tc_linear_fwd(long M, int K, int N, const float *A, int lda, const float *a_scale,
                  const float *a_shift, int a_relu, const float *W, const float *bias, float *Y,
                  double *stats, float *ws, size_t ws_bytes, bool image_ready, const pn2_bn_finalize *fin,
                  cudaStream_t st) {
    if (!tc_shape_ok(M, K, N) || ws == nullptr || ws_bytes < tc_image_bytes(K, N))
        return PN2_EUNSUPPORTED;
    // all column blocks in one launch; Bt(n,k) = W[k*N + n]
    const int nch = (N + TC_NCHUNK - 1) / TC_NCHUNK;
    return tc::run_chunk(M, K, nch == 1 ? N : TC_NCHUNK, nch, N, A, lda, a_scale, a_shift, a_relu, W, 1, N,
                         bias, Y, N, stats, stats ? stats + N : nullptr, ws, image_ready, fin, st);
}

int tc_linear_dgrad(long M, int K, int N, const float *dY, const float *W, float *dX, int ldx,
                    float *ws, size_t ws_bytes, bool image_ready, cudaStream_t st) {
    // dX[M,K] = dY[M,N] * W[K,N]^T : contraction over N, output columns = K ; Bt(k,n) = W[k*N + n]
    if (!tc_shape_ok(M, N, K) || ws == nullptr || ws_bytes < tc_image_bytes(N, K))
        return PN2_EUNSUPPORTED;
    const int nch = (K + TC_NCHUNK - 1) / TC_NCHUNK;
    return tc::run_chunk(M, N, nch == 1 ? K : TC_NCHUNK, nch, K, dY, N, nullptr, nullptr, 0, W, N, 1, nullptr,
                         dX, ldx, nullptr, nullptr, ws, image_ready, nullptr, st);
}

}  // namespace pn2

// debug hook (not in the public header): read and reset the cycle accounting of CTA 0
extern "C" __attribute__((visibility("default"))) int pn2_debug_tc_trace(long long *out16) {
#ifdef PN2_TRACE
    long long zeros[16] = {0};
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, pn2::tc::g_tc_trace, sizeof(zeros)) != cudaSuccess) return -2;
    if (cudaMemcpyToSymbol(pn2::tc::g_tc_trace, zeros, sizeof(zeros)) != cudaSuccess) return -2;
    return 0;
#else
    (void)out16;
    return -3;  // built without -DPN2_TRACE: the role loops carry no cycle accounting
#endif
}
