// pn2_api.cu -- ABI bookkeeping: version, error strings, per-device attribute cache.
#include <string.h>

#include "pn2_common.cuh"

namespace pn2 {

static thread_local char g_last_err[256] = {0};

void set_last_cuda_error(const char *msg) {
    strncpy(g_last_err, msg ? msg : "", sizeof(g_last_err) - 1);
    g_last_err[sizeof(g_last_err) - 1] = 0;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = kNumSMsB200;
        cached[dev] = v;
    }
    return cached[dev];
}

}  // namespace pn2

PN2_API int pn2_abi_version(void) { return 1; }

PN2_API const char *pn2_strerror(int code) {
    switch (code) {
        case PN2_OK: return "ok";
        case PN2_EINVAL: return "invalid argument (shape / attribute)";
        case PN2_ELAUNCH: return "CUDA launch failure";
        case PN2_EUNSUPPORTED: return "size not supported by the sm_100a kernels";
        case PN2_ENULL: return "required pointer is NULL";
        default: return "unknown pn2 error";
    }
}

PN2_API const char *pn2_last_cuda_error(void) { return pn2::g_last_err; }
