// errs.h -- error plumbing shared by the C-ABI translation units
#pragma once
#include <cuda_runtime.h>
int dsocp_set_err(int code, const char* fmt, ...);
int dsocp_require_device();
#define set_err dsocp_set_err
#define require_device dsocp_require_device
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return dsocp_set_err(e_ == cudaErrorMemoryAllocation ? DOTSOCP_ENOMEM : DOTSOCP_ECUDA, "%s:%d %s: %s", __FILE__, \
                                 __LINE__, #call, cudaGetErrorString(e_));                                    \
    } while (0)
