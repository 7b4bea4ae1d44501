// pn2_common.cuh -- shared helpers for the sm_100a kernels behind include/pn2_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pn2_b200.h"

#define PN2_API extern "C" __attribute__((visibility("default")))

namespace pn2 {

constexpr int kNumSMsB200 = 148;

// Remember the CUDA error text of a failing launch for pn2_last_cuda_error().
void set_last_cuda_error(const char *msg);

inline int finish_launch() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_last_cuda_error(cudaGetErrorString(e));
        cudaGetLastError();  // clear the sticky-less error so later calls start clean
        return PN2_ELAUNCH;
    }
    return PN2_OK;
}

inline int cuda_status(cudaError_t e) {
    if (e != cudaSuccess) {
        set_last_cuda_error(cudaGetErrorString(e));
        cudaGetLastError();
        return PN2_ELAUNCH;
    }
    return PN2_OK;
}

inline cudaStream_t as_stream(pn2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached, immutable).
int num_sms();

// Opt a kernel in to `bytes` of dynamic shared memory (or set another function attribute) ONCE per
// (kernel, device): the driver call leaves the hot path after the first launch.  The cache only grows
// (a larger request re-issues the call); immutable afterwards, so re-entrant callers see a stable value.
int opt_in_attr(const void *kernel, cudaFuncAttribute attr, int value);
template <typename K>
inline int opt_in_dyn_smem(K kernel, size_t bytes) {
    return opt_in_attr(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)bytes);
}

// ---- programmatic dependent launch -------------------------------------------------------------------------
// Every kernel of this library is launched with the programmatic-stream-serialization attribute and starts with
// pdl_enter() = `griddepcontrol.wait`: it may become resident as soon as the CTAs of the PREVIOUS kernel of the
// stream have exited (instead of after that kernel's completion and memory flush), and is held at the wait until
// that kernel has completed and its writes are visible.  What overlaps is launch latency, CTA scheduling and
// whatever a kernel does before its wait (barrier init, TMEM allocation in the tensor-core GEMMs) -- never a
// global-memory access: no kernel reads or writes global memory ahead of its wait
// (tests/test_host_logic.py::test_every_kernel_waits_for_its_predecessor).  Under stream capture the attribute
// becomes a programmatic edge of the graph.  PN2_PDL=0 launches everything fully serialised (A/B switch).
// An EARLY trigger (`griddepcontrol.launch_dependents` at kernel entry, -DPN2_PDL_EARLY_TRIGGER) measured 2 %
// SLOWER per training step than no PDL at all: the next kernel's CTAs are placed while the SMs still hold CTAs of
// the current one, wherever there is room, and the grid-stride / persistent kernels then run unbalanced
// (profiles/README_r02.md, call S).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() {
#ifdef PN2_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// at the END of a kernel's work (the tcgen05 GEMMs, before their TMEM release / last-CTA finalize): the next kernel's
// launch and prologue overlap this tail; its CTAs land on SMs that are about to be free, so the placement stays even
// (same box: 3.277 -> 3.268 ms per step; -DPN2_PDL_NO_LATE_TRIGGER builds without it)
__device__ __forceinline__ void pdl_trigger_late() {
#ifndef PN2_PDL_NO_LATE_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_trigger();
    pdl_wait();
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in finish_launch()
}
#endif

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

// Squared distance exactly as the nvcc-compiled reference kernels evaluate it
// (FMUL dy,dy ; FFMA dx,dx ; FFMA dz,dz -- see oracle/pn2_oracle.c).  Explicit
// intrinsics so no compiler flag can change the rounding sequence.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

}  // namespace pn2

#define PN2_REQUIRE(cond)                \
    do {                                 \
        if (!(cond)) return PN2_EINVAL;  \
    } while (0)
#define PN2_REQUIRE_PTR(p)              \
    do {                                \
        if ((p) == nullptr) return PN2_ENULL; \
    } while (0)
