// common.cu -- error slot, launch counter, device info.
#include "sb_common.h"
#include <string.h>

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void sb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void sb_count_launch(void) { ++g_launches; }

extern "C" const char* sb_last_error(void) { return g_err; }
extern "C" int sb_version(void) { return 100; }
extern "C" int64_t sb_launch_count(void) { return g_launches; }

extern "C" int sb_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes) {
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (sm_count) SB_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) SB_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) SB_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (smem_optin_bytes) SB_CUDA(cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return SB_OK;
}
