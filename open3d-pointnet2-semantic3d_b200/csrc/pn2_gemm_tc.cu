// pn2_gemm_tc.cu -- tcgen05 (5th-gen tensor core) 3xTF32 GEMM path.  Placeholder until the
// kernel lands: reports "unsupported" so that pn2_linear_* fall back to the exact fp32 kernel.
#include "pn2_common.cuh"

namespace pn2 {
int tc_linear_fwd(long, int, int, const float *, int, const float *, const float *, int,
                  const float *, const float *, float *, double *, cudaStream_t) {
    return PN2_EUNSUPPORTED;
}
int tc_linear_dgrad(long, int, int, const float *, const float *, float *, int, cudaStream_t) {
    return PN2_EUNSUPPORTED;
}
}  // namespace pn2
