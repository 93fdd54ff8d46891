// api.cu -- error plumbing and device queries of the C ABI (include/rlg_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

namespace rlg {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        cached_dev = dev;
        cached = n;
    }
    return cached;
}

}  // namespace rlg

extern "C" {

int rlg_version(void) { return RLG_ABI_VERSION; }

const char *rlg_last_error(void) { return rlg::g_err; }

int rlg_device_sm_count(void) {
    int n = rlg::sm_count();
    if (n < 0) return rlg::fail(-100, "rlg_device_sm_count: no CUDA device");
    return n;
}

}  // extern "C"
