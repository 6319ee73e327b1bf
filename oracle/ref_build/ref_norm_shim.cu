// TEST INFRASTRUCTURE, not product code.
// The reference's own fused residual + RMSNorm kernel (src/turbomind/kernels/norm/rms_norm.cu:287-363,
// invokeBiasResidualRMSNorm -- what UnifiedDecoder runs between the GEMMs of a layer, unified_decoder.cc:278-337 on one
// GPU) compiled in place into oracle/_ref/libtm_ref_norm.so, so tests can hold ours against it bit for bit.
#include "src/turbomind/core/logger.h"
#include "src/turbomind/kernels/norm/rms_norm.h"

#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string>

namespace turbomind::core {
Logger::Logger() = default;
Logger& Logger::Instance()
{
    static thread_local Logger* inst = new Logger();
    return *inst;
}
void Logger::LogFatalImpl(const char* file, int line, std::string message)
{
    fprintf(stderr, "[ref-norm-shim fatal] %s:%d %s\n", file ? file : "?", line, message.c_str());
    abort();
}
void Logger::LogFatalTraced(const char* file, int line, std::string message)
{
    LogFatalImpl(file, line, std::move(message));
}
void Logger::Enqueue(Level, const char*, int, std::string) {}
}  // namespace turbomind::core

extern "C" {

// residual[num, dims] += hidden[num, dims] (+ bias[dims]);  hidden <- RMSNorm(residual) * weights   (both in place)
__attribute__((visibility("default"))) int ref_bias_residual_rmsnorm(void* residual, void* hidden, const void* weights,
                                                                     const void* bias, int dims, int num, float eps, int dtype,
                                                                     void* stream)
{
    using namespace turbomind;
    if (dtype == 0) {
        invokeBiasResidualRMSNorm((half*)residual, (half*)hidden, (const half*)weights, (const half*)bias, dims, num, eps, false,
                                  (cudaStream_t)stream);
    }
    else {
        invokeBiasResidualRMSNorm((nv_bfloat16*)residual, (nv_bfloat16*)hidden, (const nv_bfloat16*)weights,
                                  (const nv_bfloat16*)bias, dims, num, eps, false, (cudaStream_t)stream);
    }
    return (int)cudaGetLastError();
}

// The reference's per-head q / k RMSNorm on fused QKV rows (rms_norm.cu:139-260, invokeQkRMSNorm; Qwen3 family), in place.
// qkv: [token_num, stride] T; heads = q_heads + k_heads (+ whatever follows in the row is not touched).
__attribute__((visibility("default"))) int ref_qk_rmsnorm(void* qkv, long stride, const void* q_weight, const void* k_weight,
                                                          int token_num, int q_heads, int k_heads, int head_dim, float eps,
                                                          int dtype, void* stream)
{
    using namespace turbomind;
    const DataType dt = dtype == 0 ? kFloat16 : kBfloat16;
    // the kernel addresses head h of token t at t * stride(0) + h * head_dim: a 3-D view over the q and k heads
    Tensor x{qkv, core::Layout{{(ssize_t)token_num, (ssize_t)(q_heads + k_heads), (ssize_t)head_dim},
                               {(ssize_t)stride, (ssize_t)head_dim, (ssize_t)1}},
             dt, kDEVICE};
    Tensor qw{const_cast<void*>(q_weight), core::Layout{{(ssize_t)head_dim}}, dt, kDEVICE};
    Tensor kw{const_cast<void*>(k_weight), core::Layout{{(ssize_t)head_dim}}, dt, kDEVICE};
    invokeQkRMSNorm(x, qw, kw, q_heads, k_heads, eps, false, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

}  // extern "C"
