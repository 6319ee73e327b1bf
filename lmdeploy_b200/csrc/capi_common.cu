#include "../../include/b200_turbomind.h"
#include "host_utils.h"
#include <cstring>

namespace b200 {
std::string& last_error_ref()
{
    static thread_local std::string s;
    return s;
}
}  // namespace b200

extern "C" const char* b200_last_error(void)
{
    return b200::last_error_ref().c_str();
}

extern "C" const char* b200_version(void)
{
    return "b200tm 0.1.0 sm_100a";
}

extern "C" int b200_ipc_alloc(size_t bytes, void** ptr, unsigned char handle[64])
{
    if (!ptr || !handle || bytes == 0) {
        return b200::fail(B200_ERR_INVALID_ARG, "ipc_alloc: bad arguments");
    }
    B200_CUDA_TRY(cudaMalloc(ptr, bytes));
    B200_CUDA_TRY(cudaMemset(*ptr, 0, bytes));
    B200_CUDA_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    B200_CUDA_TRY(cudaIpcGetMemHandle(&h, *ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle, &h, 64);
    return 0;
}

extern "C" int b200_ipc_open(const unsigned char handle[64], void** ptr)
{
    if (!ptr || !handle) {
        return b200::fail(B200_ERR_INVALID_ARG, "ipc_open: bad arguments");
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    B200_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int b200_ipc_close(void* ptr)
{
    B200_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int b200_ipc_free(void* ptr)
{
    B200_CUDA_TRY(cudaFree(ptr));
    return 0;
}
