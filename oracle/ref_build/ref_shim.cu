// TEST INFRASTRUCTURE, not product code.
// Thin C-ABI shim around the REFERENCE's own decode attention (compiled in place from /root/reference by
// oracle/ref_build/Makefile into oracle/_ref/libtm_ref_attention.so; no reference source is copied into this repo).
// It lets tests/ and bench.py run the unmodified TurboMind kernels
//   dispatchDecoding<T>(AttentionParams<T>)    src/turbomind/kernels/attention/decoding.cu:13-39
//   invokeProcessKV_v2                         src/turbomind/kernels/attention/kv_cache_utils_v2.cu:18-210
// on the same inputs as ours.  The argument block below carries exactly what
// UnifiedAttentionLayer::core_attention fills in (models/llama/unified_attention_layer.cc:470-609) and
// init_rope_kernel_param (models/attention_weight.cc:37-93).
#include "src/turbomind/core/logger.h"
#include "src/turbomind/kernels/attention/attention_params.h"
#include "src/turbomind/kernels/attention/decoding.h"
#include "src/turbomind/kernels/attention/kv_cache_utils_v2.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string>

// ---- the reference logger pulls in a queue library that is not in the image; these three members are all the
// ---- attention path references (core/check.cc, core/scope.cc) ------------------------------------------------
namespace turbomind::core {
Logger::Logger() = default;
Logger& Logger::Instance()
{
    static thread_local Logger* inst = new Logger();
    return *inst;
}
void Logger::LogFatalImpl(const char* file, int line, std::string message)
{
    fprintf(stderr, "[ref-shim fatal] %s:%d %s\n", file ? file : "?", line, message.c_str());
    abort();
}
void Logger::Enqueue(Level, const char* file, int line, std::string message)
{
    fprintf(stderr, "[ref-shim] %s:%d %s\n", file ? file : "?", line, message.c_str());
}
}  // namespace turbomind::core

extern "C" {

struct ref_rope {
    int   type;  // RopeType: 0 null, 1 default, 2 linear, 3 dynamic, 4 yarn, 5 llama3
    int   dim;
    float base;
    float factor;
    int   max_position_embeddings;
    float yarn_attention_factor, yarn_beta_fast, yarn_beta_slow;
    float llama3_low_freq_factor, llama3_high_freq_factor;
    int   llama3_original_max_position_embeddings;
};

struct ref_attn_args {
    void*       out;
    void*       q;
    void*       k;
    void*       v;
    long long   stride;
    void*       q_bias;
    void*       k_bias;
    void*       v_bias;
    const int*  cu_q_len;
    const int*  cu_k_len;
    const int*  readonly_block_num;
    const bool* finished;
    char**      block_ptrs;
    const int*  cu_block_nums;
    int         layer_offset;
    int         block_len;
    int         token_num;
    int         batch_size;
    int         max_q_len;
    int         max_k_len;
    int         num_heads;
    int         num_kv_heads;
    int         size_per_head;
    float       inv_sqrt_dh;
    ref_rope    rope;
    int         quant_policy;
    int         max_split_k;  // <= 0: the engine's default min(max(1, 4096 / token_num), 128)
    int*        split_cnt;    // [4096]
    float*      partial_O;    // [4096 * num_heads * size_per_head]
    float*      partial_ML;   // [4096 * num_heads * 2]
    int         dtype;        // 0 f16, 1 bf16
    void*       stream;
};

}  // extern "C"

namespace {

using namespace turbomind;

RopeKernelParam make_rope(const ref_rope& r)
{
    // init_rope_kernel_param, models/attention_weight.cc:37-93
    RopeKernelParam k{};
    k.type         = static_cast<RopeType>(r.type);
    k.dim          = r.dim;
    k.scale_factor = -std::log2(r.base) / r.dim;
    k.inv_factor   = (k.type == RopeType::kDynamic) ? 1.f : ((r.factor != 0.f) ? 1.0 / r.factor : 1.f);
    if (k.type == RopeType::kYarn) {
        const double PI   = 3.14159265358979323846;
        auto         corr = [&](float n) {
            return (r.dim * std::log(r.max_position_embeddings / (n * 2 * PI))) / (2 * std::log(r.base));
        };
        float low  = std::max((float)std::floor(corr(r.yarn_beta_fast)), 0.f);
        float high = std::min((float)std::ceil(corr(r.yarn_beta_slow)), r.dim - 1.f);
        if (low == high) {
            high += 0.001f;
        }
        k.yarn.ramp_inv_factor_div_2   = 1.0 / (high - low) / 2.0;
        k.yarn.ramp_inv_factor_mul_min = 1.0 / (high - low) * low;
        k.yarn.attention_factor        = r.yarn_attention_factor;
    }
    else if (k.type == RopeType::kLlama3) {
        float inv_diff = 1.0 / (r.llama3_high_freq_factor - r.llama3_low_freq_factor);
        k.llama3.alpha = r.llama3_original_max_position_embeddings / (2 * 3.14159265358979323846) * inv_diff;
        k.llama3.beta  = r.llama3_low_freq_factor * inv_diff;
    }
    k.mrope_mode = MropeMode::kNone;
    return k;
}

template<class T>
AttentionParams<T> make_params(const ref_attn_args& a)
{
    AttentionParams<T> p{};
    p.out = (T*)a.out, p.q = (T*)a.q, p.k = (T*)a.k, p.v = (T*)a.v, p.stride = a.stride;
    p.q_bias = (T*)a.q_bias, p.k_bias = (T*)a.k_bias, p.v_bias = (T*)a.v_bias;
    p.cu_q_len = a.cu_q_len, p.cu_k_len = a.cu_k_len, p.readonly_block_num = a.readonly_block_num;
    p.finished          = a.finished;
    p.block_iter_params = BlockIteratorParams{a.block_ptrs, a.cu_block_nums, a.layer_offset, a.block_len};
    p.token_num = a.token_num, p.batch_size = a.batch_size, p.max_q_len = a.max_q_len, p.max_k_len = a.max_k_len;
    p.num_heads = a.num_heads, p.num_kv_heads = a.num_kv_heads, p.size_per_head = a.size_per_head;
    p.inv_sqrt_dh   = a.inv_sqrt_dh;
    p.sinks         = nullptr;
    p.scale_sinks   = 0.f;
    p.window_size   = 256 << 20;  // unified_attention_layer.cc:552-555
    p.rope_param    = make_rope(a.rope);
    p.use_logn_attn = false;
    p.max_position_embeddings = a.rope.max_position_embeddings;
    p.split_cnt = a.split_cnt, p.partial_ML = a.partial_ML, p.partial_O = a.partial_O;
    p.max_split_k = a.max_split_k > 0 ? a.max_split_k : std::min(std::max(1, 4096 / a.token_num), 128);
    int dev = 0;
    cudaGetDevice(&dev);
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    p.arch         = major * 100 + minor * 10;
    p.stream       = (cudaStream_t)a.stream;
    p.quant_policy = a.quant_policy;
    return p;
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int ref_decode_attention(const ref_attn_args* a)
{
    if (a->dtype == 0) {
        dispatchDecoding(make_params<half>(*a));
    }
    else {
        dispatchDecoding(make_params<nv_bfloat16>(*a));
    }
    return (int)cudaGetLastError();
}

// prefill-side KV writer: bias + RoPE + quantise + page store for token_num new tokens
__attribute__((visibility("default"))) int ref_process_kv(const ref_attn_args* a)
{
    auto run = [&](auto tag) {
        using T = decltype(tag);
        auto p  = make_params<T>(*a);
        invokeProcessKV_v2_(p);
    };
    if (a->dtype == 0) {
        run(half{});
    }
    else {
        run(nv_bfloat16{});
    }
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) const char* ref_version(void)
{
    return "lmdeploy 0.16.0 src/turbomind/kernels/attention (decoding_sm80_128, reduce, kv_cache_utils_v2), sm_100a build";
}
}
