// Library-level C-ABI helpers: version, last-error string, device properties.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "../../include/dsr_b200.h"

static thread_local char g_err[512] = "";

void dsr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int dsr_check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        dsr_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return DSR_ERR_CUDA;
    }
    return DSR_OK;
}

int dsr_num_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

extern "C" int dsr_version(void) { return DSR_B200_VERSION; }
extern "C" const char* dsr_last_error_string(void) { return g_err; }
extern "C" int dsr_device_sm_count(void) { return dsr_num_sms(); }
extern "C" int dsr_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}
