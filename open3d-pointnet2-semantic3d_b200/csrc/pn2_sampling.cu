// pn2_sampling.cu -- farthest point sampling + gather_point(+grad) for sm_100a.
//
// Replaces tf_ops/tf_sampling.cu:111-206 of the reference (not a translation):
//
// FPS: one persistent CTA per cloud.  The cloud is staged once (coalesced AoS read ->
// SoA shared memory), every thread then keeps its points AND their running minimum
// distance in registers for all m-1 rounds -- the reference re-reads a (32,n) global
// scratch row and most of the cloud every round.  The per-round arg-max is two
// `redux.sync` instructions per level (max over the distance bit pattern, then min over a
// tie key among the lanes holding the max) with ONE __syncthreads per round (double-buffered
// per-warp slots) instead of the reference's 9-level shared-memory tree with 18 barriers.
//
// Bit-exactness contract: distance = fma(dz,dz,fma(dx,dx,dy*dy)), running
// min by fminf, strict '>' arg-max whose ties resolve to the lowest (k mod 512) then lowest
// k -- exactly the order the reference's 512-thread strided scan + left-biased tree gives.
#include <stdlib.h>

#include <cooperative_groups.h>
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_reduce.cuh>

#include "pn2_common.cuh"

namespace pn2 {

// Packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2): two independent IEEE round-to-nearest operations per
// instruction, so results are bit-identical to the scalar forms -- the FPS distance update issues half the FP
// instructions.  A pair lives in one 64-bit register.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2_t fadd2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fmul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// sqdist_ref for two points at once: fma(dz,dz, fma(dx,dx, dy*dy)) with d = p + (-centroid)
__device__ __forceinline__ f32x2_t sqdist_ref2(f32x2_t x, f32x2_t y, f32x2_t z, f32x2_t nx, f32x2_t ny, f32x2_t nz) {
    const f32x2_t dx = fadd2(x, nx), dy = fadd2(y, ny), dz = fadd2(z, nz);
    return ffma2(dz, dz, ffma2(dx, dx, fmul2(dy, dy)));
}


// ---- tie key: smaller wins.  (k mod 512) major, (k div 512) minor ----------------------
__device__ __forceinline__ unsigned tie_key(int k) {
    return ((unsigned)(k & 511) << 22) | (unsigned)(k >> 9);
}
__device__ __forceinline__ int key_to_k(unsigned key) {
    return (int)(((key & 0x3FFFFFu) << 9) | (key >> 22));
}

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_reg_kernel(int n, int m, const float *__restrict__ inp, int *__restrict__ out) {
    pdl_enter();
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(smem_raw);  // [2][32]
    float *xs = reinterpret_cast<float *>(slots + 64);
    const int npad = THREADS * PPT;
    float *ys = xs + npad;
    float *zs = ys + npad;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = inp + (size_t)blockIdx.x * n * 3;
    int *dst = out + (size_t)blockIdx.x * m;

    // stage the cloud: coalesced AoS global read -> SoA shared memory
    for (int e = t; e < n * 3; e += THREADS) {
        float v = __ldg(src + e);
        int k = e / 3, c = e - k * 3;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    __syncthreads();

    float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int k = t + i * THREADS;
        bool ok = k < n;
        px[i] = ok ? xs[k] : 0.f;
        py[i] = ok ? ys[k] : 0.f;
        pz[i] = ok ? zs[k] : 0.f;
        pd[i] = ok ? 1e38f : -1.f;  // padding lanes can never exceed best (= -1, strict >)
    }

    int old = 0;
    if (t == 0) dst[0] = 0;
    // visiting order inside a thread must follow the tie order: classes of (k mod 512)
    constexpr int R = (THREADS >= 512) ? 1 : 512 / THREADS;
    for (int j = 1; j < m; ++j) {
        const float x1 = xs[old], y1 = ys[old], z1 = zs[old];
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int i = r; i < PPT; i += R) {
                float d = sqdist_ref(px[i] - x1, py[i] - y1, pz[i] - z1);
                float d2 = fminf(d, pd[i]);
                pd[i] = d2;
                if (d2 > best) {
                    best = d2;
                    besti = i;
                }
            }
        }
        const bool has = best >= 0.f;
        unsigned db = has ? __float_as_uint(best) : 0u;
        unsigned key = has ? tie_key(t + besti * THREADS) : 0xFFFFFFFFu;
        unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, db);
        unsigned wkey = __reduce_min_sync(0xFFFFFFFFu, db == wmax ? key : 0xFFFFFFFFu);
        unsigned long long *sl = slots + (j & 1) * 32;
        if (lane == 0) sl[warp] = ((unsigned long long)wmax << 32) | wkey;
        __syncthreads();
        unsigned long long v = lane < NW ? sl[lane] : 0x00000000FFFFFFFFull;
        unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
        unsigned gmax = __reduce_max_sync(0xFFFFFFFFu, d2);
        unsigned gkey = __reduce_min_sync(0xFFFFFFFFu, d2 == gmax ? k2 : 0xFFFFFFFFu);
        old = gkey == 0xFFFFFFFFu ? 0 : key_to_k(gkey);
        if (t == 0) dst[j] = old;
    }
}

// Spatially pruned variant for large clouds (one CTA of 1024 threads, 8 points per thread).
// The points are sorted once by a 30-bit Morton code (cub::BlockRadixSort, blocked arrangement), so
// that the 256 points a warp keeps in registers form a compact box.  In a round a warp first
// evaluates the distance formula on its BOX (coordinate differences clamped to the box: a lower
// bound of every point's distance, exact in floating point because subtraction, multiplication,
// fma and their roundings are monotonic); if that bound is not below the warp's largest running
// minimum, no running minimum in the warp can change and the warp re-posts its cached candidate.
// After the first few dozen rounds ~85% of the warps skip.  Results are bit-identical to the plain
// kernel: the same arithmetic on the points that do change, and the tie order is carried by
// explicit keys (each thread visits its points in ascending key order, so strict '>' keeps the
// lowest key).
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_pruned_kernel(int n, int m, const float *__restrict__ inp, int *__restrict__ out) {
    pdl_enter();
    constexpr int NW = THREADS / 32;
    typedef cub::BlockRadixSort<unsigned, THREADS, PPT, int> Sort;
    typedef cub::BlockReduce<float, THREADS> Reduce;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(smem_raw);  // [2][32]
    float *bb = reinterpret_cast<float *>(slots + 64);                               // [8] cloud box
    float *xs = bb + 8;
    const int npad = THREADS * PPT;
    float *ys = xs + npad;
    float *zs = ys + npad;
    typename Sort::TempStorage &sort_tmp = *reinterpret_cast<typename Sort::TempStorage *>(zs + npad);
    typename Reduce::TempStorage &red_tmp = *reinterpret_cast<typename Reduce::TempStorage *>(zs + npad);

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = inp + (size_t)blockIdx.x * n * 3;
    int *dst = out + (size_t)blockIdx.x * m;

    for (int e = t; e < n * 3; e += THREADS) {
        float v = __ldg(src + e);
        int k = e / 3, c = e - k * 3;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    __syncthreads();

    // cloud bounding box (6 block reductions, once)
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = t; k < n; k += THREADS) {
        lo[0] = fminf(lo[0], xs[k]); hi[0] = fmaxf(hi[0], xs[k]);
        lo[1] = fminf(lo[1], ys[k]); hi[1] = fmaxf(hi[1], ys[k]);
        lo[2] = fminf(lo[2], zs[k]); hi[2] = fmaxf(hi[2], zs[k]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = Reduce(red_tmp).Reduce(lo[a], cub::Min());
        if (t == 0) bb[a] = v;
        __syncthreads();
        v = Reduce(red_tmp).Reduce(hi[a], cub::Max());
        if (t == 0) bb[3 + a] = v;
        __syncthreads();
    }
    float org[3], inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        org[a] = bb[a];
        const float ext = bb[3 + a] - bb[a];
        inv[a] = ext > 0.f ? 1023.0f / ext : 0.f;
    }

    // Morton keys of a blocked slice, block-wide sort (key, original index)
    unsigned keys[PPT];
    int vals[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = t * PPT + i;
        unsigned code = 0xFFFFFFFFu;  // padding sorts last
        if (k < n) {
            unsigned q[3];
            q[0] = (unsigned)fminf(fmaxf((xs[k] - org[0]) * inv[0], 0.f), 1023.f);
            q[1] = (unsigned)fminf(fmaxf((ys[k] - org[1]) * inv[1], 0.f), 1023.f);
            q[2] = (unsigned)fminf(fmaxf((zs[k] - org[2]) * inv[2], 0.f), 1023.f);
            code = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                unsigned v = q[a] & 1023u;  // spread 10 bits to every third position
                v = (v | (v << 16)) & 0x030000FFu;
                v = (v | (v << 8)) & 0x0300F00Fu;
                v = (v | (v << 4)) & 0x030C30C3u;
                v = (v | (v << 2)) & 0x09249249u;
                code |= v << a;
            }
        }
        keys[i] = code;
        vals[i] = k;
    }
    __syncthreads();
    Sort(sort_tmp).Sort(keys, vals);
    __syncthreads();

    // registers: coordinates, running minimum, tie key; a thread's points in ascending key order
    float px[PPT], py[PPT], pz[PPT], pd[PPT];
    unsigned tk[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = vals[i];
        const bool ok = keys[i] != 0xFFFFFFFFu && k < n;
        px[i] = ok ? xs[k] : 0.f;
        py[i] = ok ? ys[k] : 0.f;
        pz[i] = ok ? zs[k] : 0.f;
        pd[i] = ok ? 1e38f : -1.f;
        tk[i] = ok ? tie_key(k) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int a = 0; a < PPT - 1; ++a) {
#pragma unroll
        for (int b = 0; b < PPT - 1 - a; ++b) {
            if (tk[b] > tk[b + 1]) {
                unsigned tu = tk[b]; tk[b] = tk[b + 1]; tk[b + 1] = tu;
                float f;
                f = px[b]; px[b] = px[b + 1]; px[b + 1] = f;
                f = py[b]; py[b] = py[b + 1]; py[b + 1] = f;
                f = pz[b]; pz[b] = pz[b + 1]; pz[b + 1] = f;
                f = pd[b]; pd[b] = pd[b + 1]; pd[b + 1] = f;
            }
        }
    }
    // warp box over the real points
    float wlo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, whi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        if (pd[i] > 0.f) {
            wlo[0] = fminf(wlo[0], px[i]); whi[0] = fmaxf(whi[0], px[i]);
            wlo[1] = fminf(wlo[1], py[i]); whi[1] = fmaxf(whi[1], py[i]);
            wlo[2] = fminf(wlo[2], pz[i]); whi[2] = fmaxf(whi[2], pz[i]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            wlo[a] = fminf(wlo[a], __shfl_xor_sync(0xFFFFFFFFu, wlo[a], o));
            whi[a] = fmaxf(whi[a], __shfl_xor_sync(0xFFFFFFFFu, whi[a], o));
        }
    }

    // the coordinates stay packed in pairs for the rounds (FADD2 / FMUL2 / FFMA2)
    f32x2_t X2[PPT / 2], Y2[PPT / 2], Z2[PPT / 2];
#pragma unroll
    for (int i = 0; i < PPT; i += 2) {
        X2[i >> 1] = pack2(px[i], px[i + 1]);
        Y2[i >> 1] = pack2(py[i], py[i + 1]);
        Z2[i >> 1] = pack2(pz[i], pz[i + 1]);
    }
    int old = 0;
    if (t == 0) dst[0] = 0;
    // cached warp candidate: FLT_MAX forces the first update of a warp that owns real points; a warp
    // of padding only never has a candidate (its box is empty and every test below says "skip")
    unsigned wmax = whi[0] >= wlo[0] ? 0x7F7FFFFFu : 0u, wkey = 0xFFFFFFFFu;
    for (int j = 1; j < m; ++j) {
        const float x1 = xs[old], y1 = ys[old], z1 = zs[old];
        const float bx = fmaxf(fmaxf(wlo[0] - x1, x1 - whi[0]), 0.f);
        const float by = fmaxf(fmaxf(wlo[1] - y1, y1 - whi[1]), 0.f);
        const float bz = fmaxf(fmaxf(wlo[2] - z1, z1 - whi[2]), 0.f);
        const float dmin = sqdist_ref(bx, by, bz);
        if (dmin < __uint_as_float(wmax)) {  // warp-uniform: some running minimum may change
            float best = -1.f;
            unsigned bkey = 0xFFFFFFFFu;
            const f32x2_t nx = pack2(-x1, -x1), ny = pack2(-y1, -y1), nz = pack2(-z1, -z1);
#pragma unroll
            for (int i = 0; i < PPT; i += 2) {
                float da, db;
                unpack2(sqdist_ref2(X2[i >> 1], Y2[i >> 1], Z2[i >> 1], nx, ny, nz), da, db);
                const float d2a = fminf(da, pd[i]);
                pd[i] = d2a;
                if (d2a > best) {
                    best = d2a;
                    bkey = tk[i];
                }
                const float d2b = fminf(db, pd[i + 1]);
                pd[i + 1] = d2b;
                if (d2b > best) {
                    best = d2b;
                    bkey = tk[i + 1];
                }
            }
            const bool has = best >= 0.f;
            unsigned db = has ? __float_as_uint(best) : 0u;
            wmax = __reduce_max_sync(0xFFFFFFFFu, db);
            wkey = __reduce_min_sync(0xFFFFFFFFu, (has && db == wmax) ? bkey : 0xFFFFFFFFu);
        }
        unsigned long long *sl = slots + (j & 1) * 32;
        if (lane == 0) sl[warp] = ((unsigned long long)wmax << 32) | wkey;
        __syncthreads();
        unsigned long long v = lane < NW ? sl[lane] : 0x00000000FFFFFFFFull;
        unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
        unsigned gmax = __reduce_max_sync(0xFFFFFFFFu, d2);
        unsigned gkey = __reduce_min_sync(0xFFFFFFFFu, d2 == gmax ? k2 : 0xFFFFFFFFu);
        old = gkey == 0xFFFFFFFFu ? 0 : key_to_k(gkey);
        if (t == 0) dst[j] = old;
    }
}

// Variant for 8192 < n <= 16384: coordinates stay in shared memory (196 KB), only the
// running minimum lives in registers.
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_smem_kernel(int n, int m, const float *__restrict__ inp, int *__restrict__ out) {
    pdl_enter();
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(smem_raw);
    float *xs = reinterpret_cast<float *>(slots + 64);
    const int npad = THREADS * PPT;
    float *ys = xs + npad;
    float *zs = ys + npad;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = inp + (size_t)blockIdx.x * n * 3;
    int *dst = out + (size_t)blockIdx.x * m;
    for (int e = t; e < npad * 3; e += THREADS) {
        int k = e / 3, c = e - k * 3;
        float v = e < n * 3 ? __ldg(src + e) : 0.f;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    __syncthreads();
    float pd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) pd[i] = (t + i * THREADS) < n ? 1e38f : -1.f;
    int old = 0;
    if (t == 0) dst[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = xs[old], y1 = ys[old], z1 = zs[old];
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            int k = t + i * THREADS;
            float d = sqdist_ref(xs[k] - x1, ys[k] - y1, zs[k] - z1);
            float d2 = fminf(d, pd[i]);
            pd[i] = d2;
            if (d2 > best) {
                best = d2;
                besti = i;
            }
        }
        const bool has = best >= 0.f;
        unsigned db = has ? __float_as_uint(best) : 0u;
        unsigned key = has ? tie_key(t + besti * THREADS) : 0xFFFFFFFFu;
        unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, db);
        unsigned wkey = __reduce_min_sync(0xFFFFFFFFu, db == wmax ? key : 0xFFFFFFFFu);
        unsigned long long *sl = slots + (j & 1) * 32;
        if (lane == 0) sl[warp] = ((unsigned long long)wmax << 32) | wkey;
        __syncthreads();
        unsigned long long v = lane < NW ? sl[lane] : 0x00000000FFFFFFFFull;
        unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
        unsigned gmax = __reduce_max_sync(0xFFFFFFFFu, d2);
        unsigned gkey = __reduce_min_sync(0xFFFFFFFFu, d2 == gmax ? k2 : 0xFFFFFFFFu);
        old = gkey == 0xFFFFFFFFu ? 0 : key_to_k(gkey);
        if (t == 0) dst[j] = old;
    }
}

// Clouds larger than one SM can hold: a thread-block CLUSTER per cloud.  CTA `rank` of the
// cluster keeps the contiguous slice [rank*1024*PPT, (rank+1)*1024*PPT) of the cloud in its shared
// memory (SoA) and the running minima of its slice in registers, for all rounds -- nothing is
// re-read from L2/HBM.  A round is: local update + arg-max (as in fps_smem_kernel), ONE
// __syncthreads, warp 0 reduces the 32 warp candidates and pushes the CTA's candidate
// {distance bits | tie key, x, y, z} into the exchange slot [round parity][rank] of EVERY CTA of the
// cluster through distributed shared memory, ONE cluster barrier, then every warp picks the winner
// among the <= 16 candidates it finds in its own shared memory (max distance, then min tie key:
// the same total order as inside a CTA, so the result is bit-identical) and takes the winner's
// coordinates from the record -- no global or remote read on the critical path.
// Slices are multiples of 1024 points, so a thread's points t + 1024*i all share (k mod 512) and are
// visited in ascending k: the strict '>' keeps the lowest tie key, as in the other kernels.
struct __align__(16) FpsCand {
    unsigned long long dk;  // (distance bits << 32) | tie key ; key 0xFFFFFFFF = no candidate
    float x, y, z, pad;
};
constexpr int kFpsMaxCluster = 16;
constexpr bool kFpsClusterDefault = true;  // verified on the B200: bit-identical at every config-5 size, 1.1-14x faster

template <int PPT>
__global__ void __launch_bounds__(1024, 1)
fps_cluster_kernel(int n, int m, const float *__restrict__ inp, int *__restrict__ out) {
    pdl_enter();
    namespace cg = cooperative_groups;
    constexpr int THREADS = 1024, SLICE = THREADS * PPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(smem_raw);      // [2][32]
    FpsCand *exch = reinterpret_cast<FpsCand *>(slots + 64);                            // [2][16]
    float *xs = reinterpret_cast<float *>(exch + 2 * kFpsMaxCluster);
    float *ys = xs + SLICE;
    float *zs = ys + SLICE;

    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / cs;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = inp + (size_t)cloud * n * 3;
    int *dst = out + (size_t)cloud * m;
    const int k0 = rank * SLICE;                       // first point of this CTA's slice
    const int cnt = max(0, min(SLICE, n - k0));        // points it really holds

    // stage the slice: coalesced AoS global read -> SoA shared memory (tail padded with zeros)
    for (int e = t; e < SLICE * 3; e += THREADS) {
        int k = e / 3, c = e - k * 3;
        float v = k < cnt ? __ldg(src + (size_t)k0 * 3 + e) : 0.f;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    float pd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) pd[i] = (t + i * THREADS) < cnt ? 1e38f : -1.f;
    float x1 = __ldg(src), y1 = __ldg(src + 1), z1 = __ldg(src + 2);  // round 0 selects point 0
    if (rank == 0 && t == 0) dst[0] = 0;
    __syncthreads();
    cluster.sync();  // every CTA of the cluster is running: its shared memory may be written

    for (int j = 1; j < m; ++j) {
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = t + i * THREADS;
            float d = sqdist_ref(xs[k] - x1, ys[k] - y1, zs[k] - z1);
            float d2 = fminf(d, pd[i]);
            pd[i] = d2;
            if (d2 > best) {
                best = d2;
                besti = i;
            }
        }
        const bool has = best >= 0.f;
        unsigned db = has ? __float_as_uint(best) : 0u;
        unsigned key = has ? tie_key(k0 + t + besti * THREADS) : 0xFFFFFFFFu;
        unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, db);
        unsigned wkey = __reduce_min_sync(0xFFFFFFFFu, db == wmax ? key : 0xFFFFFFFFu);
        const int par = j & 1;
        if (lane == 0) slots[par * 32 + warp] = ((unsigned long long)wmax << 32) | wkey;
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = slots[par * 32 + lane];  // NW == 32: one slot per lane
            unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
            unsigned cmax = __reduce_max_sync(0xFFFFFFFFu, d2);
            unsigned ckey = __reduce_min_sync(0xFFFFFFFFu, d2 == cmax ? k2 : 0xFFFFFFFFu);
            if (lane < cs) {
                FpsCand c;
                c.dk = ((unsigned long long)cmax << 32) | ckey;
                const int kl = ckey == 0xFFFFFFFFu ? 0 : key_to_k(ckey) - k0;  // local index
                c.x = xs[kl];
                c.y = ys[kl];
                c.z = zs[kl];
                c.pad = 0.f;
                FpsCand *peer = cluster.map_shared_rank(exch, lane);  // CTA `lane`'s exchange area
                peer[par * kFpsMaxCluster + rank] = c;
            }
            __syncwarp();
        }
        cluster.sync();  // release the pushes / acquire everybody's
        // winner among the cs candidates (same order: max distance bits, then min tie key)
        unsigned long long v = lane < cs ? exch[par * kFpsMaxCluster + lane].dk : 0x00000000FFFFFFFFull;
        unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
        unsigned gmax = __reduce_max_sync(0xFFFFFFFFu, d2);
        unsigned gkey = __reduce_min_sync(0xFFFFFFFFu, d2 == gmax ? k2 : 0xFFFFFFFFu);
        const unsigned who = __ballot_sync(0xFFFFFFFFu, lane < cs && d2 == gmax && k2 == gkey);
        const FpsCand *w = exch + par * kFpsMaxCluster + (who ? __ffs(who) - 1 : 0);
        x1 = w->x;
        y1 = w->y;
        z1 = w->z;
        if (rank == 0 && t == 0) dst[j] = gkey == 0xFFFFFFFFu ? 0 : key_to_k(gkey);
    }
}

// EXPERIMENTAL (never run on a GPU yet; only reachable through pn2_fps_cluster_mb / the
// PN2_EXPERIMENTAL tests): same data flow as fps_cluster_kernel, but the per-round cluster barrier
// -- measured at ~0.9 us (UCGABAR + MEMBAR.ALL.GPU + CCTL.IVALL) -- is replaced by a point-to-point
// handshake: after pushing its candidate into a peer's exchange slot, the pushing lane arrives
// (release, cluster scope) on THAT peer's mbarrier of the round's parity; every thread then waits
// (acquire, cluster scope) on its own CTA's mbarrier for the `cs` arrivals of the round.  Reuse is
// safe without any further synchronisation: a peer can push round j+2 (same parity as j) only after
// it has seen this CTA's round-(j+1) push, which warp 0 issues after the __syncthreads that every
// warp reaches only when it has finished reading round j's records.
__device__ __forceinline__ unsigned smem_addr_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_remote_arrive_release(void *local_bar, unsigned peer_rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_addr_u32(local_bar)), "r"(peer_rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(void *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_LOOP_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}

template <int PPT>
__global__ void __launch_bounds__(1024, 1)
fps_cluster_mb_kernel(int n, int m, const float *__restrict__ inp, int *__restrict__ out) {
    pdl_enter();
    namespace cg = cooperative_groups;
    constexpr int THREADS = 1024, SLICE = THREADS * PPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(smem_raw);      // [2][32]
    FpsCand *exch = reinterpret_cast<FpsCand *>(slots + 64);                            // [2][16]
    unsigned long long *xbar = reinterpret_cast<unsigned long long *>(exch + 2 * kFpsMaxCluster);  // [2]
    float *xs = reinterpret_cast<float *>(xbar + 2);
    float *ys = xs + SLICE;
    float *zs = ys + SLICE;

    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / cs;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = inp + (size_t)cloud * n * 3;
    int *dst = out + (size_t)cloud * m;
    const int k0 = rank * SLICE;
    const int cnt = max(0, min(SLICE, n - k0));

    if (t == 0) {
        mbar_init(&xbar[0], (unsigned)cs);
        mbar_init(&xbar[1], (unsigned)cs);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = t; e < SLICE * 3; e += THREADS) {
        int k = e / 3, c = e - k * 3;
        float v = k < cnt ? __ldg(src + (size_t)k0 * 3 + e) : 0.f;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    float pd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) pd[i] = (t + i * THREADS) < cnt ? 1e38f : -1.f;
    float x1 = __ldg(src), y1 = __ldg(src + 1), z1 = __ldg(src + 2);
    if (rank == 0 && t == 0) dst[0] = 0;
    __syncthreads();
    cluster.sync();  // every CTA runs and its mbarriers are initialised

    for (int j = 1; j < m; ++j) {
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = t + i * THREADS;
            float d = sqdist_ref(xs[k] - x1, ys[k] - y1, zs[k] - z1);
            float d2 = fminf(d, pd[i]);
            pd[i] = d2;
            if (d2 > best) {
                best = d2;
                besti = i;
            }
        }
        const bool has = best >= 0.f;
        unsigned db = has ? __float_as_uint(best) : 0u;
        unsigned key = has ? tie_key(k0 + t + besti * THREADS) : 0xFFFFFFFFu;
        unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, db);
        unsigned wkey = __reduce_min_sync(0xFFFFFFFFu, db == wmax ? key : 0xFFFFFFFFu);
        const int par = j & 1;
        if (lane == 0) slots[par * 32 + warp] = ((unsigned long long)wmax << 32) | wkey;
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = slots[par * 32 + lane];
            unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
            unsigned cmax = __reduce_max_sync(0xFFFFFFFFu, d2);
            unsigned ckey = __reduce_min_sync(0xFFFFFFFFu, d2 == cmax ? k2 : 0xFFFFFFFFu);
            if (lane < cs) {
                FpsCand c;
                c.dk = ((unsigned long long)cmax << 32) | ckey;
                const int kl = ckey == 0xFFFFFFFFu ? 0 : key_to_k(ckey) - k0;
                c.x = xs[kl];
                c.y = ys[kl];
                c.z = zs[kl];
                c.pad = 0.f;
                FpsCand *peer = cluster.map_shared_rank(exch, lane);
                peer[par * kFpsMaxCluster + rank] = c;
                mbar_remote_arrive_release(&xbar[par], (unsigned)lane);  // orders the store above
            }
            __syncwarp();
        }
        // barrier `par` serves rounds par, par+2, ...: round j is its ((j-1)>>1)-th use (j = 1, 2 -> first use,
        // phase parity 0; j = 3, 4 -> second use, parity 1; ...)
        mbar_wait_acquire_cluster(&xbar[par], (unsigned)(((j - 1) >> 1) & 1));
        unsigned long long v = lane < cs ? exch[par * kFpsMaxCluster + lane].dk : 0x00000000FFFFFFFFull;
        unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
        unsigned gmax = __reduce_max_sync(0xFFFFFFFFu, d2);
        unsigned gkey = __reduce_min_sync(0xFFFFFFFFu, d2 == gmax ? k2 : 0xFFFFFFFFu);
        const unsigned who = __ballot_sync(0xFFFFFFFFu, lane < cs && d2 == gmax && k2 == gkey);
        const FpsCand *w = exch + par * kFpsMaxCluster + (who ? __ffs(who) - 1 : 0);
        x1 = w->x;
        y1 = w->y;
        z1 = w->z;
        if (rank == 0 && t == 0) dst[j] = gkey == 0xFFFFFFFFu ? 0 : key_to_k(gkey);
    }
    cluster.sync();  // no CTA leaves while a peer could still push into it
}

// Fallback for clouds larger than one SM can hold: one CTA per cloud, coordinates read
// through L1/L2 each round, running minimum in a caller-provided (b,n) global scratch.
// Same arithmetic and tie order; used only beyond 16384 points (sweep sizes of config 5).
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(int n, int m, const float *__restrict__ inp, float *__restrict__ temp,
                  int *__restrict__ out) {
    pdl_enter();
    constexpr int THREADS = 1024, NW = 32;
    __shared__ unsigned long long slots[64];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = inp + (size_t)blockIdx.x * n * 3;
    float *td = temp + (size_t)blockIdx.x * n;
    int *dst = out + (size_t)blockIdx.x * m;
    for (int k = t; k < n; k += THREADS) td[k] = 1e38f;
    int old = 0;
    if (t == 0) dst[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(src + old * 3), y1 = __ldg(src + old * 3 + 1),
                    z1 = __ldg(src + old * 3 + 2);
        float best = -1.f;
        int bestk = 0;
        for (int k = t; k < n; k += THREADS) {
            float d = sqdist_ref(__ldg(src + k * 3) - x1, __ldg(src + k * 3 + 1) - y1,
                                 __ldg(src + k * 3 + 2) - z1);
            float d2 = fminf(d, td[k]);
            td[k] = d2;
            if (d2 > best) {
                best = d2;
                bestk = k;
            }
        }
        const bool has = best >= 0.f;
        unsigned db = has ? __float_as_uint(best) : 0u;
        unsigned key = has ? tie_key(bestk) : 0xFFFFFFFFu;
        unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, db);
        unsigned wkey = __reduce_min_sync(0xFFFFFFFFu, db == wmax ? key : 0xFFFFFFFFu);
        unsigned long long *sl = slots + (j & 1) * 32;
        if (lane == 0) sl[warp] = ((unsigned long long)wmax << 32) | wkey;
        __syncthreads();
        unsigned long long v = lane < NW ? sl[lane] : 0x00000000FFFFFFFFull;
        unsigned d2 = (unsigned)(v >> 32), k2 = (unsigned)v;
        unsigned gmax = __reduce_max_sync(0xFFFFFFFFu, d2);
        unsigned gkey = __reduce_min_sync(0xFFFFFFFFu, d2 == gmax ? k2 : 0xFFFFFFFFu);
        old = gkey == 0xFFFFFFFFu ? 0 : key_to_k(gkey);
        if (t == 0) dst[j] = old;
    }
}

template <int THREADS, int PPT>
static int launch_fps_reg(int b, int n, int m, const float *inp, int *out, cudaStream_t st) {
    size_t smem = 64 * sizeof(unsigned long long) + (size_t)THREADS * PPT * 3 * sizeof(float);
    auto kern = fps_reg_kernel<THREADS, PPT>;
    if (smem > 48 * 1024) {
        int rc = opt_in_dyn_smem(kern, smem);
        if (rc) return rc;
    }
    launch_k(kern, b, THREADS, smem, st, n, m, inp, out);
    return finish_launch();
}

template <int THREADS, int PPT>
static int launch_fps_pruned(int b, int n, int m, const float *inp, int *out, cudaStream_t st) {
    typedef cub::BlockRadixSort<unsigned, THREADS, PPT, int> Sort;
    typedef cub::BlockReduce<float, THREADS> Reduce;
    size_t tmp = sizeof(typename Sort::TempStorage) > sizeof(typename Reduce::TempStorage)
                     ? sizeof(typename Sort::TempStorage) : sizeof(typename Reduce::TempStorage);
    size_t smem = 64 * sizeof(unsigned long long) + 8 * sizeof(float) +
                  (size_t)THREADS * PPT * 3 * sizeof(float) + tmp + 16;
    auto kern = fps_pruned_kernel<THREADS, PPT>;
    int rc = opt_in_dyn_smem(kern, smem);
    if (rc) return rc;
    launch_k(kern, b, THREADS, smem, st, n, m, inp, out);
    return finish_launch();
}

template <int THREADS, int PPT>
static int launch_fps_smem(int b, int n, int m, const float *inp, int *out, cudaStream_t st) {
    size_t smem = 64 * sizeof(unsigned long long) + (size_t)THREADS * PPT * 3 * sizeof(float);
    auto kern = fps_smem_kernel<THREADS, PPT>;
    int rc = opt_in_dyn_smem(kern, smem);
    if (rc) return rc;
    launch_k(kern, b, THREADS, smem, st, n, m, inp, out);
    return finish_launch();
}

template <int PPT, bool HANDSHAKE = false>
static int launch_fps_cluster(int b, int cs, int n, int m, const float *inp, int *out,
                              cudaStream_t st) {
    size_t smem = 64 * sizeof(unsigned long long) + 2 * kFpsMaxCluster * sizeof(FpsCand) +
                  (HANDSHAKE ? 2 * sizeof(unsigned long long) : 0) +
                  (size_t)1024 * PPT * 3 * sizeof(float);
    auto kern = HANDSHAKE ? fps_cluster_mb_kernel<PPT> : fps_cluster_kernel<PPT>;
    int rc = opt_in_dyn_smem(kern, smem);
    if (rc) return rc;
    if (cs > 8) {
        rc = opt_in_attr(reinterpret_cast<const void *>(kern), cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (rc) return rc;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * cs), 1, 1);
    cfg.blockDim = dim3(1024, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
        cudaGetLastError();
        return PN2_EUNSUPPORTED;  // this cluster shape cannot be co-scheduled on the device
    }
    if (pdl_enabled()) cfg.numAttrs = 2;
    rc = cuda_status(cudaLaunchKernelEx(&cfg, kern, n, m, inp, out));
    if (rc) return rc;
    return finish_launch();
}

// Cluster shape: the fewest points per thread (shortest round) whose cluster fits the limit of 16
// CTAs and, when possible, lets all b clusters run in one wave; a shape the device cannot
// co-schedule (e.g. no GPC with 16 free SMs) falls through to the next larger slice per CTA.
// PN2_EUNSUPPORTED when no shape works (n > 262144, or clusters unavailable).
template <bool HANDSHAKE = false>
static int dispatch_fps_cluster(int b, int n, int m, const float *inp, int *out, cudaStream_t st) {
    for (int p = 2; p <= 16; p *= 2) {
        int need = ceil_div(n, 1024 * p), cs = 2;
        while (cs < need) cs *= 2;
        if (cs > kFpsMaxCluster) continue;
        if ((long)b * cs > num_sms() && p < 16) continue;
        int rc = p == 2   ? launch_fps_cluster<2, HANDSHAKE>(b, cs, n, m, inp, out, st)
                 : p == 4 ? launch_fps_cluster<4, HANDSHAKE>(b, cs, n, m, inp, out, st)
                 : p == 8 ? launch_fps_cluster<8, HANDSHAKE>(b, cs, n, m, inp, out, st)
                          : launch_fps_cluster<16, HANDSHAKE>(b, cs, n, m, inp, out, st);
        if (rc != PN2_EUNSUPPORTED) return rc;
    }
    return PN2_EUNSUPPORTED;
}

// ---- gather_point / grad ---------------------------------------------------------------
__global__ void gather_point_kernel(int n, int m, long total, const float *__restrict__ inp,
                                    const int *__restrict__ idx, float *__restrict__ out) {
    pdl_enter();
    // one thread per output float: e = (cloud*m + j)*3 + c ; coalesced writes
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / 3;
        int c = (int)(e - row * 3);
        long cloud = row / m;
        int a = __ldg(idx + row);
        out[e] = __ldg(inp + (cloud * n + a) * 3 + c);
    }
}

__global__ void gather_point_grad_kernel(int n, int m, long total,
                                         const float *__restrict__ out_g,
                                         const int *__restrict__ idx,
                                         float *__restrict__ inp_g) {
    pdl_enter();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
         e += (long)gridDim.x * blockDim.x) {
        long row = e / 3;
        int c = (int)(e - row * 3);
        long cloud = row / m;
        int a = __ldg(idx + row);
        atomicAdd(inp_g + (cloud * n + a) * 3 + c, __ldg(out_g + e));
    }
}

// ---- prob_sample: prefix sum in the reference's rounding order + inverse-CDF search ------
// fp32 addition is not associative, so bit-exact parity with tf_sampling.cu:7-92 fixes the
// addition DAG (see oracle/pn2_oracle.c cumsum_row_ref): chunks of 8192 elements; local quad
// prefixes; balanced tree sums S over the quad totals; inclusive prefixes
// P(p) = S(p) + P(p - lowbit(p+1)); a compensated carry between chunks.  The DAG is evaluated
// here without any shared-memory tree: one CTA per row, 1024 threads, thread t owns the quad
// pair (2t, 2t+1); the first tree level is a register add, levels 1-5 are warp shuffles, levels
// 6-10 are shuffles of one warp over the 32 warp totals, and the down-sweep runs the same
// ladder backwards.  Positions beyond the end of a ragged chunk hold zeros; in-range positions
// never read them (every read goes to a lower position), so no bounds test is needed.
constexpr int kScanThreads = 1024;
constexpr int kScanChunk = 8192;  // part of the rounding sequence (tf_sampling.cu:9,15), not a tuning knob

__global__ void __launch_bounds__(kScanThreads)
prob_cdf_kernel(int n, const float *__restrict__ inp, float *__restrict__ out) {
    pdl_enter();
    __shared__ float warp_tot[32];
    __shared__ float chunk_total;
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const float *row = inp + (size_t)blockIdx.x * n;
    float *orow = out + (size_t)blockIdx.x * n;
    float run = 0.0f, comp = 0.0f;

    for (int j = 0; j < n; j += kScanChunk) {
        const int len = min(n - j, kScanChunk);
        const int nq = (len + 3) >> 2;
        float e[2][4];
        float qt[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 4 * (2 * t + h);
            if (k + 3 < len) {
                const float a = __ldg(row + j + k), b = __ldg(row + j + k + 1);
                const float c = __ldg(row + j + k + 2), d = __ldg(row + j + k + 3);
                const float ba = __fadd_rn(b, a), dc = __fadd_rn(d, c);
                e[h][0] = a;
                e[h][1] = ba;
                e[h][2] = __fadd_rn(c, ba);
                e[h][3] = __fadd_rn(dc, ba);
                qt[h] = e[h][3];
            } else {  // ragged last quad: left to right from zero; beyond the chunk: zeros
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (k + i < len) acc = __fadd_rn(acc, __ldg(row + j + k + i));
                    e[h][i] = acc;
                }
                qt[h] = acc;
            }
        }
        // up-sweep: a = S(block of lowbit(lane+1) pairs ending at this pair)
        float a = __fadd_rn(qt[1], qt[0]);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float o = __shfl_up_sync(full, a, d);
            if (((lane + 1) & (2 * d - 1)) == 0) a = __fadd_rn(a, o);
        }
        if (lane == 31) warp_tot[w] = a;
        __syncthreads();
        if (w == 0) {  // tree over the 32 warp totals, then their inclusive prefixes
            float x = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float o = __shfl_up_sync(full, x, d);
                if (((lane + 1) & (2 * d - 1)) == 0) x = __fadd_rn(x, o);
            }
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) {
                const float o = __shfl_up_sync(full, x, d);
                if (((lane + 1) & (2 * d - 1)) == d && lane + 1 > d) x = __fadd_rn(x, o);
            }
            warp_tot[lane] = x;
        }
        __syncthreads();
        // down-sweep inside the warp: P(p) = S(p) + P(p - lowbit(p+1)); the pair in front of the
        // warp's first block is the last pair of the previous warp (its prefix is warp_tot[w-1])
        const float before = w > 0 ? warp_tot[w - 1] : 0.0f;
        if (lane == 31) a = warp_tot[w];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float o = __shfl_up_sync(full, a, d);
            if (((lane + 1) & (2 * d - 1)) == d) {
                if (lane + 1 > d) a = __fadd_rn(a, o);
                else if (w > 0) a = __fadd_rn(a, before);
            }
        }
        float pprev = __shfl_up_sync(full, a, 1);  // inclusive prefix of pair t-1
        if (lane == 0) pprev = before;
        const bool has_prev = t > 0;
        const float pre0 = has_prev ? __fadd_rn(qt[0], pprev) : qt[0];  // prefix of quad 2t
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 4 * (2 * t + h);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (k + i < len) {
                    float v = e[h][i];
                    if (h == 1) v = __fadd_rn(v, pre0);
                    else if (has_prev) v = __fadd_rn(v, pprev);
                    orow[j + k + i] = __fadd_rn(v, run);
                }
            }
        }
        if (t == ((nq - 1) >> 1)) chunk_total = ((nq - 1) & 1) ? a : pre0;
        __syncthreads();
        const float tt = __fadd_rn(chunk_total, comp);
        const float r2 = __fadd_rn(run, tt);
        comp = __fsub_rn(tt, __fsub_rn(r2, run));
        run = r2;
    }
}

// tf_sampling.cu:94-110: r = n-1; for k = base, base/2, .. 1: if (r >= k && cdf[r-k] >= q) r -= k.
// The CDF need not be monotone in fp32 (tree-ordered sum), so this exact search -- not a generic
// lower bound -- is what defines the result.  One thread per draw; the top levels of the search
// hit the same few CDF entries for every thread and stay in L1.
__global__ void prob_search_kernel(int b, int n, int m, int base, const float *__restrict__ cdf,
                                   const float *__restrict__ query, int *__restrict__ result) {
    pdl_enter();
    for (int i = blockIdx.y; i < b; i += gridDim.y) {
        const float *c = cdf + (size_t)i * n;
        const float total = __ldg(c + n - 1);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
            const float q = __fmul_rn(__ldg(query + (size_t)i * m + j), total);
            int r = n - 1;
            for (int k = base; k >= 1; k >>= 1)
                if (r >= k && __ldg(c + r - k) >= q) r -= k;
            result[(size_t)i * m + j] = r;
        }
    }
}

}  // namespace pn2

using namespace pn2;

PN2_API int pn2_fps(int b, int n, int m, const float *inp, float *temp, int *out,
                    pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m > 0);  // tf_sampling.cpp:121-123 "expects positive npoint"
    if (b == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp);
    PN2_REQUIRE_PTR(out);
    cudaStream_t st = as_stream(s);
    if (n <= 128) return launch_fps_reg<128, 1>(b, n, m, inp, out, st);
    if (n <= 256) return launch_fps_reg<128, 2>(b, n, m, inp, out, st);
    if (n <= 512) return launch_fps_reg<256, 2>(b, n, m, inp, out, st);
    if (n <= 1024) return launch_fps_reg<256, 4>(b, n, m, inp, out, st);
    if (n <= 2048) return launch_fps_reg<512, 4>(b, n, m, inp, out, st);
    // large clouds: spatial pruning pays once a round costs more than the box tests
    // (PN2_FPS_PRUNE=0 selects the plain register kernel)
    static const bool prune = !(getenv("PN2_FPS_PRUNE") && getenv("PN2_FPS_PRUNE")[0] == '0');
    if (n <= 4096) return prune && n > 2048 && m > 64 ? launch_fps_pruned<1024, 4>(b, n, m, inp, out, st)
                                                      : launch_fps_reg<1024, 4>(b, n, m, inp, out, st);
    // 4096 < n <= 8192: 512 threads x 16 points (r02 A/B on B200: 0.457 vs 0.516 ms at 16 x 8192 -> 1024 with 1024 x 8,
    // bit-identical: half the warps at the per-round barrier and in the final arg-max outweigh the coarser pruning
    // boxes).  PN2_FPS_T=1024 | 256 selects the other shapes for A/B runs.
    static const int fps_t = getenv("PN2_FPS_T") ? atoi(getenv("PN2_FPS_T")) : 512;
    if (n <= 8192) {
        if (!(prune && m > 64)) return launch_fps_reg<1024, 8>(b, n, m, inp, out, st);
        if (fps_t == 1024) return launch_fps_pruned<1024, 8>(b, n, m, inp, out, st);
        if (fps_t == 256) return launch_fps_pruned<256, 32>(b, n, m, inp, out, st);
        return launch_fps_pruned<512, 16>(b, n, m, inp, out, st);
    }
    // more than 8192 points: a thread-block cluster per cloud keeps the whole cloud on chip
    // (PN2_FPS_CLUSTER=0/1 overrides the default)
    static const char *cl_env = getenv("PN2_FPS_CLUSTER");
    static const bool use_cluster = cl_env ? cl_env[0] != '0' : kFpsClusterDefault;
    if (use_cluster) {
        // per-round exchange: push + remote mbarrier arrive (3-15 % faster than the cluster barrier on
        // B200 at 16 k-262 k points, profiles/README_r02.md); PN2_FPS_HANDSHAKE=0 selects the barrier
        static const bool handshake = !(getenv("PN2_FPS_HANDSHAKE") && getenv("PN2_FPS_HANDSHAKE")[0] == '0');
        int rc = handshake ? dispatch_fps_cluster<true>(b, n, m, inp, out, st)
                           : dispatch_fps_cluster<false>(b, n, m, inp, out, st);
        if (rc != PN2_EUNSUPPORTED) return rc;
    }
    if (n <= 16384) return launch_fps_smem<1024, 16>(b, n, m, inp, out, st);
    // beyond what a cluster holds (or no cluster available): streaming kernel; it needs the
    // (b,n) scratch the reference also requires (tf_sampling.cpp:143-146 allocates (32,n))
    if (temp == nullptr) return PN2_ENULL;
    launch_k(fps_stream_kernel, b, 1024, 0, st, n, m, inp, temp, out);
    return finish_launch();
}

PN2_API int pn2_fps_cluster(int b, int n, int m, const float *inp, int *out, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m > 0);
    if (b == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp);
    PN2_REQUIRE_PTR(out);
    return dispatch_fps_cluster<false>(b, n, m, inp, out, as_stream(s));
}

PN2_API int pn2_fps_cluster_mb(int b, int n, int m, const float *inp, int *out, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m > 0);
    if (b == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp);
    PN2_REQUIRE_PTR(out);
    return dispatch_fps_cluster<true>(b, n, m, inp, out, as_stream(s));
}

PN2_API int pn2_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out,
                             pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    long total = (long)b * m * 3;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(out);
    int threads = 256;
    long blocks = ceil_div<long>(total, threads);
    if (blocks > 148L * 16) blocks = 148L * 16;
    launch_k(gather_point_kernel, (int)blocks, threads, 0, as_stream(s), n, m, total, inp, idx, out);
    return finish_launch();
}

PN2_API int pn2_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx,
                                  float *inp_g, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    if (b == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp_g);
    cudaStream_t st = as_stream(s);
    int rc = cuda_status(cudaMemsetAsync(inp_g, 0, sizeof(float) * (size_t)b * n * 3, st));
    if (rc) return rc;
    long total = (long)b * m * 3;
    if (total == 0) return PN2_OK;
    PN2_REQUIRE_PTR(out_g);
    PN2_REQUIRE_PTR(idx);
    int threads = 256;
    long blocks = ceil_div<long>(total, threads);
    if (blocks > 148L * 16) blocks = 148L * 16;
    launch_k(gather_point_grad_kernel, (int)blocks, threads, 0, st, n, m, total, out_g, idx, inp_g);
    return finish_launch();
}

PN2_API int pn2_cumsum(int b, int n, const float *inp, float *out, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0);
    if (b == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp);
    PN2_REQUIRE_PTR(out);
    launch_k(prob_cdf_kernel, b, kScanThreads, 0, as_stream(s), n, inp, out);
    return finish_launch();
}

PN2_API int pn2_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r,
                            float *temp, int *out, pn2_stream_t s) {
    PN2_REQUIRE(b >= 0 && n > 0 && m >= 0);
    if (b == 0 || m == 0) return PN2_OK;
    PN2_REQUIRE_PTR(inp_p);
    PN2_REQUIRE_PTR(inp_r);
    PN2_REQUIRE_PTR(temp);  // (b,n) floats: the CDF, as in the reference (tf_sampling.cpp:104-108)
    PN2_REQUIRE_PTR(out);
    cudaStream_t st = as_stream(s);
    launch_k(prob_cdf_kernel, b, kScanThreads, 0, st, n, inp_p, temp);
    int rc = finish_launch();
    if (rc) return rc;
    int base = 1;
    while (base < n) base <<= 1;
    const int threads = 256;
    dim3 grid((unsigned)min(ceil_div(m, threads), 148 * 8), (unsigned)min(b, 65535));
    launch_k(prob_search_kernel, grid, threads, 0, st, b, n, m, base, temp, inp_r, out);
    return finish_launch();
}
