// pn2_api.cu -- ABI bookkeeping: version, error strings, per-device attribute cache.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "pn2_common.cuh"

namespace pn2 {

static thread_local char g_last_err[256] = {0};

void set_last_cuda_error(const char *msg) {
    strncpy(g_last_err, msg ? msg : "", sizeof(g_last_err) - 1);
    g_last_err[sizeof(g_last_err) - 1] = 0;
}

static std::atomic<int> g_sm_budget{0};

static int device_sms() {
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

// SMs the persistent kernels size their grids for: all of them, or the budget set by pn2_set_sm_budget
int num_sms() {
    const int all = device_sms(), b = g_sm_budget.load(std::memory_order_relaxed);
    return b > 0 && b < all ? b : all;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("PN2_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

namespace {
struct AttrEntry {
    const void *kernel;
    int attr, dev, value;
};
constexpr int kMaxAttr = 256;
AttrEntry g_attr[kMaxAttr];
int g_attr_n = 0;
std::mutex g_attr_mu;
}  // namespace

int opt_in_attr(const void *kernel, cudaFuncAttribute attr, int value) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    std::lock_guard<std::mutex> lock(g_attr_mu);
    AttrEntry *hit = nullptr;
    for (int i = 0; i < g_attr_n; ++i)
        if (g_attr[i].kernel == kernel && g_attr[i].attr == (int)attr && g_attr[i].dev == dev) {
            hit = &g_attr[i];
            break;
        }
    if (hit && hit->value >= value) return PN2_OK;
    int rc = cuda_status(cudaFuncSetAttribute(kernel, attr, value));
    if (rc) return rc;
    if (hit) hit->value = value;
    else if (g_attr_n < kMaxAttr) g_attr[g_attr_n++] = AttrEntry{kernel, (int)attr, dev, value};
    return PN2_OK;
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

PN2_API int pn2_set_sm_budget(int sms) {
    if (sms < 0) return PN2_EINVAL;
    const int prev = pn2::g_sm_budget.exchange(sms, std::memory_order_relaxed);
    (void)prev;
    return PN2_OK;
}
PN2_API int pn2_get_sm_budget(void) { return pn2::num_sms(); }
