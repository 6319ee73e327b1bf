// RMSNorm family, embedding lookup, greedy argmax -- the HBM-trivial glue of the decode step.
//
// Replaces  invokeRMSNorm / invokeBiasResidualRMSNorm   (src/turbomind/kernels/norm/rms_norm.cu:21-86, 287-363,
//           rms_norm_utils.cuh:6-15), invokeEmbeddingLookup (kernels/gpt_kernels.cu:43) and the top-k=1 path of
//           Generation::Forward (generation/generation.cc:247-293, sampling.cc:113-180).
// B200 twist: the normalised activations can be written directly in the UMMA K-major tile layout
// [dim/8][MT][8] that the next GEMM's TMA bulk copy consumes (no repack kernel, no extra HBM round trip).
#include "../../include/b200_turbomind.h"
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace norm {


template<class T>
__device__ __forceinline__ void store_norm(T* out, int out_mt, int tok, int dim, int i, const T (&v)[8])
{
    // i: first of 8 consecutive dims
    if (out_mt > 0) {
        *reinterpret_cast<uint4*>(out + ((size_t)(i >> 3) * out_mt + tok) * 8) = *reinterpret_cast<const uint4*>(&v[0]);
    }
    else {
        *reinterpret_cast<uint4*>(out + (size_t)tok * dim + i) = *reinterpret_cast<const uint4*>(&v[0]);
    }
}

__device__ __forceinline__ float block_sum(float v)
{
    __shared__ float s_part[32];
    __shared__ float s_total;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0) {
        s_part[threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_part[i];
        s_total = t;
    }
    __syncthreads();
    return s_total;
}

// residual (optional): r = residual + hidden (+ bias) in T, stored back; then y = T(float(r) * inv_rms) * w.
// Single pass: every thread keeps its (<= kMaxChunks x 8) elements in registers between the sum-of-squares and the
// normalisation, so the kernel is one global round trip + one block reduction.
constexpr int kMaxChunks = 4;

template<class T, bool RESIDUAL>
__global__ void __launch_bounds__(1024) rmsnorm_kernel(T* out, int out_mt, T* residual, const T* hidden, const T* w,
                                                       const T* bias, int dim, float eps)
{
    pdl_launch_dependents();
    pdl_wait();
    using Ops     = TypeOps<T>;
    const int tok = blockIdx.x;
    T*        r   = residual + (size_t)tok * dim;
    T         rv[kMaxChunks][8];
    float     acc = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        const int i = (c * blockDim.x + threadIdx.x) * 8;
        if (i < dim) {
            *reinterpret_cast<uint4*>(&rv[c][0]) = *reinterpret_cast<const uint4*>(r + i);
            if constexpr (RESIDUAL) {
                T hv[8];
                *reinterpret_cast<uint4*>(&hv[0]) = *reinterpret_cast<const uint4*>(hidden + (size_t)tok * dim + i);
#pragma unroll
                for (int j = 0; j < 8; ++j) rv[c][j] = Ops::add(rv[c][j], hv[j]);
                if (bias) {
                    T bv[8];
                    *reinterpret_cast<uint4*>(&bv[0]) = *reinterpret_cast<const uint4*>(bias + i);
#pragma unroll
                    for (int j = 0; j < 8; ++j) rv[c][j] = Ops::add(rv[c][j], bv[j]);
                }
                *reinterpret_cast<uint4*>(r + i) = *reinterpret_cast<const uint4*>(&rv[c][0]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float f = Ops::to_f32(rv[c][j]);
                acc           = fmaf(f, f, acc);
            }
        }
    }
    const float sum     = block_sum(acc);
    const float inv_rms = rsqrtf(sum / (float)dim + eps);
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        const int i = (c * blockDim.x + threadIdx.x) * 8;
        if (i < dim) {
            T wv[8], ov[8];
            *reinterpret_cast<uint4*>(&wv[0]) = *reinterpret_cast<const uint4*>(w + i);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ov[j] = Ops::mul(Ops::from_f32(Ops::to_f32(rv[c][j]) * inv_rms), wv[j]);
            }
            store_norm(out, out_mt, tok, dim, i, ov);
        }
    }
}

template<class T>
__global__ void embedding_kernel(T* out, const T* table, const int* ids, int dim, int vocab)
{
    pdl_launch_dependents();
    pdl_wait();
    const int tok = blockIdx.x;
    int       id  = ids[tok];
    id            = min(max(id, 0), vocab - 1);
    for (int i = threadIdx.x * 8; i < dim; i += blockDim.x * 8) {
        *reinterpret_cast<uint4*>(out + (size_t)tok * dim + i) =
            *reinterpret_cast<const uint4*>(table + (size_t)id * dim + i);
    }
}

// greedy: index of the maximum of fp32-cast logits; ties -> smallest index.  Two stages: (rows x kArgChunks) CTAs
// scan a vocab slice each, then one warp per row merges the partials (deterministic).
constexpr int kArgChunks  = 16;
constexpr int kArgMaxRows = 1024;
__device__ float g_arg_val[kArgMaxRows * kArgChunks];
__device__ int   g_arg_idx[kArgMaxRows * kArgChunks];

__device__ __forceinline__ void arg_better(float& bv, int& bi, float v, int i)
{
    if (v > bv || (v == bv && i < bi)) {
        bv = v;
        bi = i;
    }
}

template<class T>
__global__ void __launch_bounds__(256) argmax_stage1_kernel(const T* logits, int vocab, int64_t ld)
{
    pdl_launch_dependents();
    pdl_wait();
    using Ops    = TypeOps<T>;
    const int row = blockIdx.x, ch = blockIdx.y;
    const int per = (vocab + kArgChunks - 1) / kArgChunks;
    const int lo = ch * per, hi = min(vocab, lo + per);
    const T*  lp = logits + (size_t)row * ld;
    float     bv = -INFINITY;
    int       bi = 0x7fffffff;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        arg_better(bv, bi, Ops::to_f32(lp[i]), i);
    }
    __shared__ float s_v[8];
    __shared__ int   s_i[8];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        arg_better(bv, bi, __shfl_xor_sync(0xffffffffu, bv, m), __shfl_xor_sync(0xffffffffu, bi, m));
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = bv;
        s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) arg_better(bv, bi, s_v[w], s_i[w]);
        g_arg_val[row * kArgChunks + ch] = bv;
        g_arg_idx[row * kArgChunks + ch] = bi;
    }
}

__global__ void argmax_stage2_kernel(int* out, int rows)
{
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) {
        return;
    }
    const int lane = threadIdx.x & 31;
    float     bv   = lane < kArgChunks ? g_arg_val[row * kArgChunks + lane] : -INFINITY;
    int       bi   = lane < kArgChunks ? g_arg_idx[row * kArgChunks + lane] : 0x7fffffff;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        arg_better(bv, bi, __shfl_xor_sync(0xffffffffu, bv, m), __shfl_xor_sync(0xffffffffu, bi, m));
    }
    if (lane == 0) {
        out[row] = bi;
    }
}

}  // namespace norm
}  // namespace b200

using namespace b200;
using namespace b200::norm;

static int norm_threads(int dim)
{
    const int chunks = dim / 8;
    int       th     = ((chunks + 31) / 32) * 32;
    if (th > 1024) {
        th = 1024;
    }
    if (chunks > th * kMaxChunks) {
        return 0;
    }
    // spread work: at least 128 threads
    return th < 128 ? 128 : th;
}

extern "C" int b200_rmsnorm(void* out, int out_packed_mt, const void* x, const void* weight, int tokens, int dim, float eps,
                            int dtype, void* stream)
{
    if (!out || !x || !weight || tokens <= 0 || dim % 8) {
        return fail(B200_ERR_INVALID_ARG, "bad rmsnorm arguments");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int    kBlock = norm_threads(dim);
    if (kBlock == 0) {
        return fail(B200_ERR_UNSUPPORTED, "rmsnorm: dim too large");
    }
    if (dtype == B200_DTYPE_F16) {
        launch_k(rmsnorm_kernel<half, false>, dim3(tokens), dim3(kBlock), 0, st, (half*)out, out_packed_mt, (half*)x, nullptr,
                                                               (const half*)weight, nullptr, dim, eps);
    }
    else {
        launch_k(rmsnorm_kernel<nv_bfloat16, false>, dim3(tokens), dim3(kBlock), 0, st, (nv_bfloat16*)out, out_packed_mt, (nv_bfloat16*)x,
                                                                      nullptr, (const nv_bfloat16*)weight, nullptr, dim,
                                                                      eps);
    }
    return check_launch("rmsnorm_kernel");
}

extern "C" int b200_residual_rmsnorm(void* out, int out_packed_mt, void* residual, const void* hidden, const void* weight,
                                     const void* bias, int tokens, int dim, float eps, int dtype, void* stream)
{
    if (!out || !residual || !hidden || !weight || tokens <= 0 || dim % 8) {
        return fail(B200_ERR_INVALID_ARG, "bad residual_rmsnorm arguments");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int    kBlock = norm_threads(dim);
    if (kBlock == 0) {
        return fail(B200_ERR_UNSUPPORTED, "residual_rmsnorm: dim too large");
    }
    if (dtype == B200_DTYPE_F16) {
        launch_k(rmsnorm_kernel<half, true>, dim3(tokens), dim3(kBlock), 0, st, (half*)out, out_packed_mt, (half*)residual,
                                                              (const half*)hidden, (const half*)weight,
                                                              (const half*)bias, dim, eps);
    }
    else {
        launch_k(rmsnorm_kernel<nv_bfloat16, true>, dim3(tokens), dim3(kBlock), 0, st, 
            (nv_bfloat16*)out, out_packed_mt, (nv_bfloat16*)residual, (const nv_bfloat16*)hidden,
            (const nv_bfloat16*)weight, (const nv_bfloat16*)bias, dim, eps);
    }
    return check_launch("residual_rmsnorm_kernel");
}

extern "C" int b200_embedding_lookup(void* out, const void* table, const int* token_ids, int tokens, int dim, int vocab,
                                     int dtype, void* stream)
{
    (void)dtype;
    if (!out || !table || !token_ids || tokens <= 0 || dim % 8) {
        return fail(B200_ERR_INVALID_ARG, "bad embedding arguments");
    }
    launch_k(embedding_kernel<uint16_t>, dim3(tokens), dim3(256), 0, (cudaStream_t)stream, (uint16_t*)out, (const uint16_t*)table,
                                                                         token_ids, dim, vocab);
    return check_launch("embedding_kernel");
}

extern "C" int b200_argmax(int* out, const void* logits, int rows, int vocab, int64_t ld, int dtype, void* stream)
{
    if (!out || !logits || rows <= 0 || vocab <= 0 || rows > kArgMaxRows) {
        return fail(B200_ERR_INVALID_ARG, "bad argmax arguments (rows <= 1024)");
    }
    cudaStream_t st = (cudaStream_t)stream;
    dim3         grid(rows, kArgChunks);
    if (dtype == B200_DTYPE_F16) {
        launch_k(argmax_stage1_kernel<half>, dim3(grid), dim3(256), 0, st, (const half*)logits, vocab, ld);
    }
    else {
        launch_k(argmax_stage1_kernel<nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const nv_bfloat16*)logits, vocab, ld);
    }
    if (int e = check_launch("argmax_stage1_kernel")) {
        return e;
    }
    launch_k(argmax_stage2_kernel, dim3((rows + 3) / 4), dim3(128), 0, st, out, rows);
    return check_launch("argmax_stage2_kernel");
}
