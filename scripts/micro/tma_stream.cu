// Micro-benchmark (diagnostic, not product): how fast can one persistent CTA per SM stream a
// row-major fp32 matrix [M x K] through shared memory in 128-row x 32-column boxes with 2-D TMA
// tensor loads (and write an equally sized matrix back with TMA tensor stores)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda
//   ./tma_stream M K R do_store consumers_read
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int c0, int c1, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}

constexpr int BM = 128, BK = 32, SLOT = BM * BK * 4;

struct P { long M; int K, KC, R, do_store, rd; float *sink; };

__global__ void __launch_bounds__(192, 1) stream_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmY, P p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ring = smem;
    unsigned char *stg = smem + (size_t)p.R * SLOT;  // 2 x 16 KB store staging
    uint64_t *full = reinterpret_cast<uint64_t *>(stg + 2 * SLOT);
    uint64_t *empty = full + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.R; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long tiles = (p.M + BM - 1) / BM;
    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x)
                for (int kc = 0; kc < p.KC; ++kc, ++it) {
                    const int s = it % p.R; const uint32_t ph = (it / p.R) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], SLOT);
                    tma_load_2d(ring + (size_t)s * SLOT, &tmA, kc * BK, (int)(tile * BM), &full[s]);
                }
        }
    } else if (warp < 4) {
        const int t = threadIdx.x, k4 = t & 7, r0 = t >> 3;
        float acc = 0.f;
        uint32_t it = 0;
        for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x)
            for (int kc = 0; kc < p.KC; ++kc, ++it) {
                const int s = it % p.R; const uint32_t ph = (it / p.R) & 1;
                mbar_wait(&full[s], ph);
                const unsigned char *slot = ring + (size_t)s * SLOT;
                float4 v[8];
                if (p.rd) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4 *>(slot + (r0 + 16 * i) * 128 + k4 * 16);
                }
                if (p.do_store) {
                    unsigned char *o = stg + (size_t)(it & 1) * SLOT;
                    // staging buffer (it&1) was last read by the store committed two chunks ago
                    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (p.rd) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 w = v[i]; w.x += 1.f;
                            *reinterpret_cast<float4 *>(o + (r0 + 16 * i) * 128 + k4 * 16) = w;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&empty[s]);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (t == 0) {
                        tma_store_2d(&tmY, kc * BK, (int)(tile * BM), o);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                    if (p.rd) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].w;
                    }
                    mbar_arrive(&empty[s]);
                }
            }
        if (t == 0 && p.do_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (acc == 123.456f) p.sink[0] = acc;
    }
}

static CUtensorMap make_map(float *base, long M, int K, int box_k, int box_m) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_m};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return tm;
}

int main(int argc, char **argv) {
    long M = argc > 1 ? atol(argv[1]) : 524288;
    int K = argc > 2 ? atoi(argv[2]) : 128;
    CK(cudaFree(0));
    float *A, *Y, *sink;
    CK(cudaMalloc(&A, M * K * 4)); CK(cudaMalloc(&Y, M * K * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(A, 0, M * K * 4));
    CUtensorMap tmA = make_map(A, M, K, BK, BM), tmY = make_map(Y, M, K, BK, BM);
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("M=%ld K=%d (%.0f MB in)\n", M, K, M * K * 4 / 1e6);
    for (int do_store = 0; do_store < 2; ++do_store)
        for (int rd = 0; rd < 2; ++rd)
            for (int R = 2; R <= 10; R += 2) {
                P p; p.M = M; p.K = K; p.KC = (K + BK - 1) / BK; p.R = R; p.do_store = do_store; p.rd = rd; p.sink = sink;
                size_t smem = (size_t)R * SLOT + 2 * SLOT + 512;
                if (smem > 227 * 1024) continue;
                for (int i = 0; i < 2; ++i) stream_kernel<<<148, 192, smem>>>(tmA, tmY, p);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                const int reps = 5;
                for (int i = 0; i < reps; ++i) stream_kernel<<<148, 192, smem>>>(tmA, tmY, p);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
                double bytes = (double)M * K * 4 * (1 + do_store);
                printf("store=%d consumers_read=%d R=%2d  %8.1f us  %7.1f GB/s\n", do_store, rd, R, ms * 1e3, bytes / ms / 1e6);
            }
    return 0;
}
