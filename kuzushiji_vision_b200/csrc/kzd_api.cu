// Library-level entry points: version, error string, device info.
#include <stdarg.h>

#include "kzd_common.cuh"

namespace kzd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace kzd

extern "C" __attribute__((visibility("default"))) int kzd_version(void) { return KZD_VERSION; }

extern "C" __attribute__((visibility("default"))) const char* kzd_last_error(void) { return kzd::g_err; }

extern "C" __attribute__((visibility("default"))) int kzd_device_info(int* sm, int* cc_major, int* cc_minor) {
    int dev = 0;
    KZD_CUDA(cudaGetDevice(&dev));
    int v = 0;
    KZD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (sm) *sm = v;
    KZD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_major) *cc_major = v;
    KZD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    if (cc_minor) *cc_minor = v;
    return 0;
}
