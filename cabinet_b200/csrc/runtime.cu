// Error reporting and device queries of the C-ABI.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void cabinet_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* cabinet_last_error(void) { return g_err; }
extern "C" int cabinet_abi_version(void) { return CABINET_ABI_VERSION; }

extern "C" int cabinet_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    CAB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    CAB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return CABINET_OK;
}
