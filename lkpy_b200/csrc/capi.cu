// capi.cu — error reporting and device queries of the C ABI (include/lkpy_b200.h).

#include <cstdarg>

#include "common.cuh"

namespace lk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

}  // namespace lk

extern "C" {

int lk_version(void) { return 100; }

const char *lk_last_error(void) { return lk::g_err; }

int lk_device_info(int *sm_count, int *cc)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        lk::set_error("no CUDA device: %s", cudaGetErrorString(e));
        return LK_ERR_NO_DEVICE;
    }
    int n = 0, major = 0, minor = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (sm_count) *sm_count = n;
    if (cc) *cc = major * 10 + minor;
    return LK_OK;
}

}  // extern "C"
