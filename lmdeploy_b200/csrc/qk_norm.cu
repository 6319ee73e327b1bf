// Per-head RMSNorm of q and k inside the fused QKV rows, before RoPE (Qwen3 family).
//
// Replaces  invokeQkRMSNorm  (src/turbomind/kernels/norm/rms_norm.cu:139-260), called on the QKV GEMM output by
//           UnifiedAttentionLayer (models/llama/unified_attention_layer.cc:739).
//
// The arithmetic follows the instruction sequence of the reference kernel as its release build compiles it
// (`cuobjdump -sass` of oracle/_ref/libtm_ref_norm.so, QkRMSNorm<T, float, 8, 128>): 16 lanes own one head, 8 elements
// each; sum of squares as an fp32 FMA chain over the lane's 8 elements in order, starting from 0; butterfly over lane
// distances 8, 4, 2, 1; r = rsqrt.approx(fma(sum, 1/dim, eps)); y = T(float(x) * r) *_T w.  The same grouping here makes
// the result bit-identical by construction, not by tolerance.
#include "../../include/b200_turbomind.h"
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace norm {

__device__ __forceinline__ float fma_ftz(float a, float b, float c)
{
    float d;
    asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ float add_ftz(float a, float b)
{
    float d;
    asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

__device__ __forceinline__ float mul_ftz_f(float a, float b)
{
    float d;
    asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

__device__ __forceinline__ float rsqrt_approx_ftz(float a)
{
    float d;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
    return d;
}

// One 16-lane group per (token, head); heads [0, q_heads) are q, [q_heads, q_heads + k_heads) are k.  head_dim == 128.
template<class T>
__global__ void __launch_bounds__(256) qk_rmsnorm_kernel(T* qkv, int64_t ld, const T* q_weight, const T* k_weight, int q_heads,
                                                         int k_heads, int token_num, float eps, float inv_dim)
{
    using Ops = TypeOps<T>;
    pdl_launch_dependents();
    pdl_wait();
    const int heads = q_heads + k_heads;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;  // (token, head)
    const int lane  = threadIdx.x & 15;
    const bool live = group < token_num * heads;
    // every lane of a warp takes part in the shuffles; dead groups compute on zeros and store nothing
    const int ti = live ? group / heads : 0;
    const int hi = live ? group % heads : 0;
    T*        p  = qkv + (int64_t)ti * ld + (int64_t)hi * 128 + lane * 8;
    uint4     raw = make_uint4(0, 0, 0, 0);
    if (live) {
        raw = *reinterpret_cast<const uint4*>(p);
    }
    const T* x = reinterpret_cast<const T*>(&raw);
    float    f[8];
    float    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = Ops::to_f32(x[i]);
        sum  = fma_ftz(f[i], f[i], sum);
    }
#pragma unroll
    for (int mask = 8; mask >= 1; mask >>= 1) {
        sum = add_ftz(sum, __shfl_xor_sync(0xffffffffu, sum, mask));
    }
    const float r = rsqrt_approx_ftz(fma_ftz(sum, inv_dim, eps));
    if (live) {
        const T* w  = (hi < q_heads ? q_weight : k_weight) + lane * 8;
        uint4    wr = __ldg(reinterpret_cast<const uint4*>(w));
        const T* wv = reinterpret_cast<const T*>(&wr);
        uint4    out;
        T*       o = reinterpret_cast<T*>(&out);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o[i] = Ops::mul(Ops::from_f32(mul_ftz_f(f[i], r)), wv[i]);
        }
        *reinterpret_cast<uint4*>(p) = out;
    }
}

}  // namespace norm
}  // namespace b200

using namespace b200;
using namespace b200::norm;

extern "C" int b200_qk_rmsnorm(void* qkv, int64_t stride, const void* q_weight, const void* k_weight, int token_num,
                               int num_heads, int num_kv_heads, int head_dim, float eps, int dtype, void* stream)
{
    if (!qkv || !q_weight || !k_weight || token_num <= 0 || num_heads <= 0 || num_kv_heads <= 0) {
        return fail(B200_ERR_INVALID_ARG, "bad qk_rmsnorm arguments");
    }
    if (head_dim != 128) {
        return fail(B200_ERR_UNSUPPORTED, "qk_rmsnorm: head_dim must be 128");
    }
    if (stride < (int64_t)(num_heads + num_kv_heads) * head_dim || stride % 8) {
        return fail(B200_ERR_INVALID_ARG, "qk_rmsnorm: stride must cover the q and k heads and keep rows 16-byte aligned");
    }
    if (dtype != B200_DTYPE_F16 && dtype != B200_DTYPE_BF16) {
        return fail(B200_ERR_INVALID_ARG, "dtype must be 0 or 1");
    }
    cudaStream_t  st     = (cudaStream_t)stream;
    const int64_t groups = (int64_t)token_num * (num_heads + num_kv_heads);
    const int     blocks = (int)((groups * 16 + 255) / 256);
    const float   inv    = 1.f / (float)head_dim;
    if (dtype == B200_DTYPE_F16) {
        launch_k(qk_rmsnorm_kernel<half>, dim3(blocks), dim3(256), 0, st, (half*)qkv, stride, (const half*)q_weight,
                 (const half*)k_weight, num_heads, num_kv_heads, token_num, eps, inv);
    }
    else {
        launch_k(qk_rmsnorm_kernel<nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (nv_bfloat16*)qkv, stride,
                 (const nv_bfloat16*)q_weight, (const nv_bfloat16*)k_weight, num_heads, num_kv_heads, token_num, eps, inv);
    }
    return check_launch("qk_rmsnorm_kernel");
}
