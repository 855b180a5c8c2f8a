// brov_internal.cuh — error reporting shared by the translation units of libbrov.so (the text behind
// brov_last_error() lives in brov_api.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/brov.h"

namespace brov {
int fail_msg(int code, const char* fmt, ...);
}
#define BROV_CUDA_TRY(expr)                                                                                   \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return brov::fail_msg(BROV_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)
