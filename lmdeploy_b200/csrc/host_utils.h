// Host-side helpers shared by the C-ABI translation units: thread-local error text, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <string>

namespace b200 {

std::string& last_error_ref();

inline int fail(int code, const char* msg)
{
    last_error_ref() = msg;
    return code;
}

inline int fail_cuda(cudaError_t e, const char* what)
{
    last_error_ref() = std::string(what) + ": " + cudaGetErrorString(e);
    return (int)e;
}

inline int check_launch(const char* kernel)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        return fail_cuda(e, kernel);
    }
    return 0;
}

#define B200_CUDA_TRY(expr)                                                                                            \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) {                                                                                       \
            return ::b200::fail_cuda(_e, #expr);                                                                       \
        }                                                                                                              \
    } while (0)

}  // namespace b200
