// Host-side helpers shared by the C-ABI translation units: thread-local error text, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <utility>
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

// Every kernel of the decode step is launched with programmatic stream serialization (PDL): the next kernel's CTAs
// may be scheduled -- and run their input-independent prologue: barrier init, TMEM allocation, weight prefetch --
// while the previous kernel drains; each kernel executes griddepcontrol.wait before it touches upstream data
// (common.cuh: pdl_wait / pdl_launch_dependents).  Captured into CUDA graphs as programmatic edges.
inline bool pdl_enabled()
{
    static const bool on = getenv("B200_NO_PDL") == nullptr;
    return on;
}

template<class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = grid;
    cfg.blockDim         = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = st;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// One-time, PER-DEVICE set-up at a call site (kernel attributes such as the dynamic shared-memory limit belong to the device's
// context: a process that drives several GPUs from several threads -- the reference's engine -- must repeat them on each).
// `slot` is a static of the call site; returns the current device, and true in *first for the first caller on that device
// (a concurrent second caller may also see true: the set-up calls are idempotent).
constexpr int kMaxDevices = 64;
struct PerDevice {
    std::atomic<int> value[kMaxDevices];   // 0 = not configured yet on that device
};
inline int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

#define B200_CUDA_TRY(expr)                                                                                            \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) {                                                                                       \
            return ::b200::fail_cuda(_e, #expr);                                                                       \
        }                                                                                                              \
    } while (0)

}  // namespace b200
