// sb_common.h -- error slot, CUDA checks and small device helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/sionna_b200.h"

// thread-local error message (defined in common.cu)
void sb_set_error(const char* fmt, ...);
void sb_count_launch(void);

#define SB_CHECK_ARG(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            sb_set_error(__VA_ARGS__);           \
            return SB_EINVAL;                    \
        }                                        \
    } while (0)

#define SB_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            sb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SB_ECUDA;                                                            \
        }                                                                               \
    } while (0)

#define SB_LAUNCH_CHECK()                                                                \
    do {                                                                                 \
        cudaError_t e_ = cudaGetLastError();                                             \
        if (e_ != cudaSuccess) {                                                         \
            sb_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SB_ECUDA;                                                             \
        }                                                                                \
        sb_count_launch();                                                               \
    } while (0)

static inline int sb_num_sms(void) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
