// TEST INFRASTRUCTURE, not product code.
// The reference's own PREFILL attention path for one chunk, in the order UnifiedAttentionLayer::core_attention runs it
// (src/turbomind/models/llama/unified_attention_layer.cc:613-638):
//   invokeProcessKV_v2_   (kernels/attention/kv_cache_utils_v2.h:41-67)    bias + RoPE + quantise + page store
//   invokeFlattenKV_v2_   (kv_cache_utils_v2.h:93-117)                     pages -> linear [H, 2, sum_k_len, D] in T
//   dispatchAttention<T>  (kernels/attention/attention.cu)                 causal attention over the linear K/V
// compiled in place into oracle/_ref/libtm_ref_prefill.so (a separate library: libtm_ref_attention.so, which the verified
// decode tests and bench.py load, is left exactly as it was).  The argument block and its translation are the decode
// shim's.
#include "ref_shim.cu"

#include "src/turbomind/kernels/attention/attention.h"

extern "C" {

// tmp_kv: device scratch of num_kv_heads * 2 * (sum_k_len + 64) * size_per_head elements of T
// (unified_attention_layer.cc:468); sum_k_len = cu_k_len[batch_size] (the caller knows it on the host).
__attribute__((visibility("default"))) int ref_prefill_attention(const ref_attn_args* a, void* tmp_kv, int sum_k_len)
{
    auto run = [&](auto tag) {
        using T = decltype(tag);
        auto p  = make_params<T>(*a);
        p.max_split_k        = 1;  // the engine disables split-kv for prefill (unified_attention_layer.cc:626-628)
        p.linear_iter_params = LinearIteratorParams{tmp_kv,                               // flattened KV
                                                    sum_k_len * a->size_per_head * 2,     // stride to the next head
                                                    sum_k_len * a->size_per_head};        // stride from K to V
        invokeProcessKV_v2_(p);
        invokeFlattenKV_v2_(p, sum_k_len);
        dispatchAttention(p);
    };
    if (a->dtype == 0) {
        run(half{});
    }
    else {
        run(nv_bfloat16{});
    }
    return (int)cudaGetLastError();
}

}  // extern "C"
