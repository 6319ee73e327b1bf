/*
 * b200_turbomind.h -- C ABI of the B200-native TurboMind decode hot path (libb200tm.so).
 *
 * Every entry point below replaces one reference interface on the decode path of
 * InternLM/lmdeploy 0.16.0 (paths relative to the reference root).  The reference has no extern "C"
 * surface (its FFI is the pybind11 module `_turbomind`, src/turbomind/python/bind.cpp:423-1004);
 * these are the plain-pointer equivalents of the *inner* operator boundary its own layers call
 * (SURVEY.md section 8b), plus an engine-level handle API that stands where
 * LanguageModel::Run / ModelRequest::Forward sit.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: all pointers are DEVICE pointers unless a name ends in `_host`; `stream` is a
 * cudaStream_t passed as void*; calls are asynchronous on that stream; every function returns 0 on
 * success or a negative b200_status / positive cudaError_t on failure (b200_last_error() gives text).
 * No torch types, no C++ types.  dtype: 0 = fp16, 1 = bf16.
 */
#ifndef B200_TURBOMIND_H
#define B200_TURBOMIND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#define B200_DTYPE_F16 0
#define B200_DTYPE_BF16 1

typedef enum {
    B200_OK              = 0,
    B200_ERR_INVALID_ARG = -1,
    B200_ERR_UNSUPPORTED = -2, /* shape / feature outside the hot-path scope (head_dim != 128, ...) */
    B200_ERR_WORKSPACE   = -3,
    B200_ERR_STATE       = -4
} b200_status;

/* Text of the last error raised on the calling thread ("" if none). */
B200_API const char* b200_last_error(void);
/* "b200tm <version> sm_100a" */
B200_API const char* b200_version(void);

/* ------------------------------------------------------------------------------------------------
 * RoPE parameters.  Replaces core::RopeConfig -> RopeKernelParam
 * (src/turbomind/models/attention_weight.h:15-35, attention_weight.cc:37-93,
 *  src/turbomind/models/llama/llama_rope.h:9-62).  type: 0 null, 1 default, 2 linear, 3 dynamic(unsupported),
 *  4 yarn, 5 llama3.  Derived kernel constants are computed inside the library exactly as
 *  init_rope_kernel_param does.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int   type;
    int   dim;
    float base;
    float factor;
    int   max_position_embeddings;
    float llama3_low_freq_factor;
    float llama3_high_freq_factor;
    int   llama3_original_max_position_embeddings;
    float yarn_attention_factor;
    float yarn_beta_fast;
    float yarn_beta_slow;
    /* 0 (default): the arithmetic of the reference's RELEASE build -- exp2f / sincosf are the hardware
     * approximations ex2.approx / sin.approx / cos.approx (--use_fast_math, kernels/attention/CMakeLists.txt:14-15),
     * so K pages are bit-identical to a TurboMind build on the same GPU.
     * 1: correctly rounded exp2 / sin / cos (reproducible off-GPU; what oracle/attention.py computes). */
    int exact_math;
} b200_rope_config;

/* ------------------------------------------------------------------------------------------------
 * Decode attention.  Replaces
 *   template<class T> void dispatchDecoding(const AttentionParams<T>&)
 *     (src/turbomind/kernels/attention/decoding.h:9-10, decoding.cu:13-39)
 * with the argument block AttentionParams<T> (attention_params.h:31-103) and BlockIteratorParams
 * (attention_params.h:20-25) flattened to C.  Field meaning is the reference's:
 *   q/k/v: the new token's fused QKV row per sequence, token-major, row stride `stride` elements;
 *   cu_q_len / cu_k_len: [batch+1] prefix sums (decode: q_len == 1, k_len includes the new token);
 *   block_ptrs[cu_block_nums[b] + i]: base pointer of the i-th KV block of sequence b (all layers in
 *   one block, KV page format of kernels/attention/block.h:124-219); layer_offset: byte offset of this
 *   layer inside a block (unified_attention_layer.cc:163-171); quant_policy 0 / 4 / 8.
 * The call (1) adds bias, applies RoPE to q and k, quantises k/v and stores them into the page at
 * position history_len (unless inside the read-only prefix), (2) runs split-KV flash decoding over the
 * paged cache, (3) merges split partials.  Unsupported reference features return B200_ERR_UNSUPPORTED:
 * window_size, attention sinks, log-n scaling, context parallel, head_dim != 128, q_len > 1.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    void*       out;
    const void* q;
    const void* k;
    const void* v;
    int64_t     stride;
    const void* q_bias;
    const void* k_bias;
    const void* v_bias;

    const int*     cu_q_len;
    const int*     cu_k_len;
    const int*     readonly_block_num; /* may be NULL */
    const uint8_t* finished;           /* may be NULL; bool per sequence */

    void**     block_ptrs;
    const int* cu_block_nums;
    int        layer_offset;
    int        block_len;

    int token_num;
    int batch_size;
    int max_k_len;

    int   num_heads;
    int   num_kv_heads;
    int   size_per_head;
    float inv_sqrt_dh; /* log2(e) * softmax_scale, as unified_attention_layer.cc:541-548 */
    int   window_size; /* must be 0 */

    b200_rope_config rope;
    int              quant_policy;
    int              dtype;

    /* workspace: b200_decode_attention_workspace_bytes() bytes, 256-byte aligned.  Stands where
     * split_cnt / partial_O / partial_ML (unified_attention_layer.cc:220-224) stand in the reference.  Like the reference's
     * GEMM barriers it must be ZERO when first used (it holds the arrival counters of the in-kernel fix-up, which every
     * launch leaves at zero again; a call that plans internally also clears them). */
    void*  workspace;
    size_t workspace_bytes;

    /* optional: a plan computed once per step by b200_decode_attention_plan (valid for every layer of
     * the step); if NULL the call plans internally. */
    const int* plan;
    /* optional: per-token (cos, sin) table from b200_rope_table (token_num x 64 float pairs), shared by all
     * layers of a step; NULL = computed inside the call. */
    const float* rope_table;

    /* > 0: write `out` in the UMMA K-major tile layout [num_heads*128/8][out_packed_mt][8] consumed by
     * b200_linear_forward_packed (fuses the repack in front of the `wo` GEMM); 0: token-major rows. */
    int out_packed_mt;

    /* profiling aid: bit 0 = skip the prologue (bias, RoPE, KV quantise + page store), bit 1 = skip the fix-up merge,
     * bit 2 = run the three-kernel form (prologue kernel, main kernel, fix-up kernel; with bits 0-1: the main kernel alone,
     * reading the rotated q an earlier bit-2 call left in the workspace), bit 3 = the fix-up inside the main kernel as well
     * (one launch; bit 3 alone also moves the prologue inside).  0 in production: three PDL-chained kernels (the fastest
     * form on B200, see decode_attention.cu). */
    int debug_flags;

    void* stream;
} b200_attention_params;

B200_API size_t b200_decode_attention_workspace_bytes(int batch_size, int num_heads, int num_kv_heads);
/* plan buffer: (batch_size + 1) ints, device memory */
B200_API int b200_decode_attention_plan(int* plan, const int* cu_k_len, const uint8_t* finished, int batch_size, void* stream);
B200_API int b200_rope_table(float* table, const b200_rope_config* rope, const int* cu_q_len, const int* cu_k_len,
                    int batch_size, int token_num, void* stream);
B200_API int b200_decode_attention(const b200_attention_params* params);

/* Prefill-side KV writer.  Replaces invokeProcessKV_v2 (kernels/attention/kv_cache_utils_v2.cu:18-210):
 * bias + RoPE on K, quantise K/V, store token positions [history_len, history_len + q_len) of every
 * sequence into the paged cache.  Same argument block; `q`, `out` are ignored. */
B200_API int b200_process_kv(const b200_attention_params* params);

/* Prefill ("context") attention of a chunk of q_len >= 1 new tokens per sequence: causal over keys
 * [0, history_len + i] for query i.  Replaces, for that chunk, invokeProcessKV_v2 + invokeFlattenKV_v2
 * (kernels/attention/kv_cache_utils_v2.cu:18-420) + dispatchAttention<T> (kernels/attention/attention.cu,
 * attention_universal.h:355-553 with the prefill mainloop): the call (1) adds bias, applies RoPE to q and k, quantises
 * k/v and stores them at positions [history_len, history_len + q_len) of the pages, (2) attends each query over the
 * pages (de-quantised on the fly) on tcgen05 tensor cores.  Same argument block as decode; cu_q_len / cu_k_len give
 * q_len and history_len + q_len per sequence; `out` is token-major [token_num, num_heads * 128] (out_packed_mt must be
 * 0); workspace = b200_prefill_attention_workspace_bytes(token_num, num_heads); `plan`, `rope_table`, `max_k_len`
 * are not used. */
B200_API size_t b200_prefill_attention_workspace_bytes(int token_num, int num_heads);
B200_API int b200_prefill_attention(const b200_attention_params* params);

/* Number of CTAs the attention kernel launches on the current device (for tests / profiling). */
B200_API int b200_decode_attention_grid(int quant_policy, int dtype);

/* ------------------------------------------------------------------------------------------------
 * Weight-only-quant linear layer.  Replaces
 *   LinearWeight (models/linear_weight.h:42-112) + LinearWeight::prepare (models/linear_weight.cc:101-324)
 *   LlamaLinear::Forward (models/llama/LlamaLinear.cu:226-273) -> gemm::Gemm::Run (kernels/gemm/gemm.h:23-41)
 * weight_format: 0 dense T [K,N]; 1 u4 AWQ (TurboMind pre-prepare layout: qweight int32 [K, N/8] with nibble i of
 *   word w = column 8w+i -- what lmdeploy/turbomind/weight_format.py:63-74 pack_u4_row produces -- plus
 *   scales T [K/g, N] and integer-valued zeros T [K/g, N], g = 128);
 *   2 fp8: weight float8_e4m3fn bytes [K, N] (or [N, K] with weight_is_transposed) and `scales` = the f32 block
 *   scales [K/128, N/128] (weight_scale_inv; zeros NULL) -- the non-sm90 path of the reference: block scales cast to
 *   T and applied per K-group, w = T(e4m3) *_T T(scale) (models/linear_weight.cc:138-150, gemm/transform.h:92-97);
 *   3 fp8 with the scales already expanded along N: `scales` f32 [K/128, N], one per (128-k group, output column) --
 *   the form the reference's kernel consumes after LinearWeight::prepare's expansion; a loader that permutes,
 *   interleaves or slices output columns (RoPE pair reorder, fused gate|up, TP) expands the checkpoint's blocks
 *   first and hands this format over.  Same arithmetic as format 2.
 * D[m, N] = x[m, K] . dequant(W) (+ bias); u4 dequant is w = fma_T(q, s, -(z*s)_T) exactly as
 * kernels/gemm/transform.h:48-74 / cast.cu:155; accumulation fp32 on tcgen05 tensor cores.
 * epilogue: 0 none; 1 gated SiLU over element-interleaved (gate, up) columns -> [m, N/2]
 *   (kernels/gemm/epilogue.h:158-177).
 * `prepared` is library-owned layout (b200_linear_prepared_bytes bytes); `bias` may be NULL.
 * Workspace must be zero-filled ONCE before its first use (split-K arrival counters live in it; the
 * kernels leave them zero).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int         input_dim;     /* K */
    int         output_dim;    /* N */
    int         weight_format; /* 0 dense T, 1 u4 (AWQ, g128), 2 fp8 e4m3 + 128x128 f32 block scales, 3 fp8 e4m3 +
                                  per-column f32 scales of 128-k groups */
    int         dtype;         /* activation / scale type T */
    int         group_size;    /* 128 (u4 quant group / fp8 scale block) */
    const void* prepared;      /* output of b200_linear_prepare */
    const void* bias;          /* [N] T or NULL */
} b200_linear_desc;

B200_API size_t b200_linear_prepared_bytes(const b200_linear_desc* desc);
B200_API int b200_linear_prepare(const b200_linear_desc* desc, const void* weight, const void* scales, const void* zeros,
                        int weight_is_transposed, void* prepared, void* stream);
B200_API size_t b200_linear_workspace_bytes(const b200_linear_desc* desc, int max_tokens);
/* x: [m, K] row-major (row stride ldx elements), y: [m, N or N/2] row-major (row stride ldy), 1 <= m <= 128 */
B200_API int b200_linear_forward(const b200_linear_desc* desc, const void* x, int64_t ldx, int m, void* y, int64_t ldy,
                        int epilogue, void* workspace, size_t workspace_bytes, void* stream);
/* Activations already in the UMMA K-major tile layout [K/8][mt][8] (mt = padded tokens, multiple of 16);
 * out_packed_mt > 0 writes the output in the same packed layout (feeds the next GEMM without a repack). */
B200_API int b200_linear_forward_packed(const b200_linear_desc* desc, const void* xp, int mt, int m, void* y, int64_t ldy,
                               int epilogue, int out_packed_mt, void* workspace, size_t workspace_bytes,
                               void* stream);
/* GEMM with the residual add + RMSNorm of unified_decoder.cc:278-287 fused into its split-K epilogue:
 *   residual[m, N] += x . W (+ bias)   (added in T);   out = T(float(residual) * inv_rms(row)) * norm_weight
 * `out` is token-major [m, N] (out_packed_mt == 0) or the packed tile layout.  Returns B200_ERR_UNSUPPORTED when the
 * launch plan of this shape has no K split (callers then run b200_linear_forward_packed + b200_residual_rmsnorm). */
B200_API int b200_linear_forward_packed_residual_rmsnorm(const b200_linear_desc* desc, const void* xp, int mt, int m,
                                                void* residual, const void* norm_weight, float eps, void* out,
                                                int out_packed_mt, void* workspace, size_t workspace_bytes,
                                                void* stream);
/* launch plan of a shape (host only; no device work): out8 = {n-tiles per CTA, K splits, W ring, X ring, TMEM A slots,
 * TMEM columns, first A column, dynamic shared memory bytes}.  A pure function of (format, K, N, padded m). */
B200_API int b200_linear_plan(const b200_linear_desc* desc, int m, int* out8);
/* Tuned launch plans -- the counterpart of the reference's GEMM tuning cache (gemm::Gemm::Export / Import,
 * kernels/gemm/gemm.h:43-45, dispatch_cache.cu).  `path`: text file of lines
 *   PLAN <weight_format> <padded tokens> <N> <K> <n-tiles per CTA> <K splits> <W ring slots> <X ring slots>
 * (ring slots 0 0 = keep the heuristic's depths; '#' starts a comment).  Every entry is validated against the budgets
 * the kernel relies on; one invalid entry rejects the file and leaves the current table unchanged.  Returns the number
 * of entries (>= 0) or a negative b200_status.  NULL clears the table.  Results stay deterministic: the plan is still a
 * pure function of (format, shape, tokens) for a given table.  Not thread-safe against concurrent forward calls. */
B200_API int b200_linear_load_plans(const char* path);
B200_API int b200_pack_activations(const void* x, int64_t ldx, int m, int k, int mt, int dtype, void* xp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Step glue.  Replace invokeRMSNorm / invokeBiasResidualRMSNorm (kernels/norm/rms_norm.cu:21-86, 287-363),
 * invokeEmbeddingLookup (kernels/gpt_kernels.cu:43) and the greedy (top_k = 1) path of Generation::Forward
 * (generation/generation.cc:247-293).  `out_packed_mt` > 0 writes the normalised activations directly in
 * the packed tile layout of b200_linear_forward_packed.
 *   rmsnorm:           out = T(float(x) * rsqrt(mean(x^2) + eps)) * w
 *   residual_rmsnorm:  residual += hidden (+ bias) in T (stored); out = rmsnorm(residual)
 *   argmax:            index of the largest fp32-cast logit per row; exact ties -> smallest index
 * ---------------------------------------------------------------------------------------------- */
B200_API int b200_rmsnorm(void* out, int out_packed_mt, const void* x, const void* weight, int tokens, int dim, float eps,
                 int dtype, void* stream);
B200_API int b200_residual_rmsnorm(void* out, int out_packed_mt, void* residual, const void* hidden, const void* weight,
                          const void* bias, int tokens, int dim, float eps, int dtype, void* stream);
B200_API int b200_embedding_lookup(void* out, const void* table, const int* token_ids, int tokens, int dim, int vocab,
                          int dtype, void* stream);
/* Per-head RMSNorm of the q and k heads inside fused QKV rows, in place, before RoPE (Qwen3 family).  Replaces
 * invokeQkRMSNorm (kernels/norm/rms_norm.cu:139-260; unified_attention_layer.cc:739).  qkv: [token_num, stride] T with a
 * row = [num_heads x head_dim | num_kv_heads x head_dim | v ...]; q_weight / k_weight: [head_dim] T.  head_dim == 128.
 * y = T(float(x) * rsqrt(mean(x^2) + eps)) *_T w, summed in the reference kernel's order (bit-identical). */
B200_API int b200_qk_rmsnorm(void* qkv, int64_t stride, const void* q_weight, const void* k_weight, int token_num,
                    int num_heads, int num_kv_heads, int head_dim, float eps, int dtype, void* stream);
B200_API int b200_argmax(int* out, const void* logits, int rows, int vocab, int64_t ld, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampling beyond greedy (do_sample requests).  Replaces LogitsProcessor::Forward (generation/logits_processor.cc:66-125:
 * repetition penalty, temperature) + Sampling::Forward (generation/sampling.cc:92-183: top-k sort filter, softmax, top-p /
 * min-p filter, invokeSampling with a per-sequence curandState_t) with ONE kernel, one CTA per sampled row.
 *   logits      [*, ld] T (the step's logits buffer); item i reads row rows[i] (rows NULL: i) and entries [0, vocab)
 *   temperature / top_k / top_p / min_p   per item (NULL: 1 / 0 = no top-k filter / 1 / 0); top_k is capped at 1024
 *   repetition_penalty (optional) with history[i] = device pointer to the history_len[i] token ids seen so far
 *   rng_states  b200_sampling_rng_state_bytes() bytes per sequence slot (curandState_t, XORWOW), seeded by
 *               b200_sampling_init_rng = curand_init(seed, 0, 0) as InitializeRandomStates; item i uses slot rng_index[i]
 *   out_uniform / out_kept (optional): the uniform number drawn and the size of the filtered candidate set (tests)
 *   workspace   b200_sampling_workspace_bytes(count, vocab) bytes
 * Candidate set and probabilities follow the reference's arithmetic; with top_k > 0 the draw walks the candidates in
 * descending-probability order like the reference, with top_k = 0 in vocabulary order (same distribution, no 92k-entry sort).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const void*       logits;
    int64_t           ld;
    int               dtype;
    int               count;
    int               vocab;
    const int*        rows;
    const float*      temperature;
    const int*        top_k;
    const float*      top_p;
    const float*      min_p;
    const float*      repetition_penalty;
    const int* const* history;
    const int*        history_len;
    void*             rng_states;
    const int*        rng_index;
    int*              out_ids;
    float*            out_uniform;
    int*              out_kept;
    void*             workspace;
    size_t            workspace_bytes;
    void*             stream;
} b200_sampling_params;

B200_API size_t b200_sampling_rng_state_bytes(void);
B200_API size_t b200_sampling_workspace_bytes(int count, int vocab);
B200_API int    b200_sampling_init_rng(void* states, const unsigned long long* seeds, const uint8_t* mask, int n, void* stream);
B200_API int    b200_sample(const b200_sampling_params* params);

/* ------------------------------------------------------------------------------------------------
 * Tensor-parallel device communicator over NVLink peer memory (one process per GPU; peers' buffers are
 * mapped with CUDA IPC by the host side and handed over as raw pointers).  Replaces
 *   DeviceCommImpl::AllreduceResidualBiasRMSnorm / AllReduceSum / AllGather
 *     (src/turbomind/comm/device_comm.h:21-143; native impl comm/cuda_ipc/fused_allreduce.cu:25-553)
 * with ONE kernel per call that reads every peer's partial GEMM output straight over NVLink (one-shot
 * "pull"), sums the ranks in rank order in T (as the reference's push/pull kernels do), adds the residual,
 * RMS-normalises and writes the next GEMM's input tile -- no separate collective, no NCCL on the path.
 *   peer_data[r]  : rank r's exchange buffer (b200_comm_buffer_bytes() bytes), peer-mapped
 *   peer_flags[r] : rank r's flag array     (b200_comm_flags_bytes() bytes, zero-initialised), peer-mapped
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200_comm b200_comm;
B200_API size_t b200_comm_buffer_bytes(int max_tokens, int hidden, int vocab_per_rank);
B200_API size_t b200_comm_flags_bytes(void);
B200_API int b200_comm_create(b200_comm** comm, int rank, int world_size, void* const* peer_data, void* const* peer_flags);
B200_API void b200_comm_destroy(b200_comm* comm);
/* this rank's exchange buffer: the GEMM that feeds a fused all-reduce must write its output here */
/* the exchange buffer has two halves used alternately by the two exchanges of a layer (no "done reading" round);
 * b200_comm_set_geometry fixes where the second half starts */
B200_API int b200_comm_set_geometry(b200_comm* comm, int max_tokens, int hidden);
B200_API void* b200_comm_local_buffer(b200_comm* comm, int half);
/* NVLS: `mc_data` = the multicast address that maps every rank's exchange buffer (same offsets on every rank, e.g. from
 * torch.distributed._symmetric_memory: rendezvous(...).multicast_ptr).  The fused all-reduce then takes the rank sum with
 * ONE multimem.ld_reduce per 16 bytes (reduced inside the NVSwitch) instead of `world` peer loads -- what the reference's
 * NVLS path does (comm/cuda_ipc/fused_allreduce.cu:169-274, multimem.cuh:14-45).  NULL = peer loads. */
B200_API int b200_comm_set_multicast(b200_comm* comm, void* mc_data);
/* hidden_partial: this rank's [tokens, dim] partial sums (T).  On return (stream order):
 *   residual += sum_ranks(hidden_partial) (+ bias);  out = RMSNorm(residual) * weight  (packed if out_packed_mt > 0) */
B200_API int b200_comm_allreduce_residual_rmsnorm(b200_comm* comm, void* out, int out_packed_mt, void* residual,
                                         const void* hidden_partial, const void* weight, const void* bias, int tokens,
                                         int dim, float eps, int dtype, void* stream);
/* Vocab-sharded greedy sampling: local argmax per rank, (value, index) pairs exchanged over peer memory,
 * every rank returns the global token id (replaces the logits all-gather of language_model.cc:306-334).
 * vocab_per_rank = row stride of the shard (padded); valid_per_rank = how many of this rank's columns are real
 * vocabulary entries -- pad columns never win (the reference bounds its sampling by vocab_size, not
 * vocab_size_padded: generation/sampling.cc:132). */
B200_API int b200_comm_argmax_allgather(b200_comm* comm, int* out_tokens, const void* local_logits, int rows,
                               int vocab_per_rank, int valid_per_rank, int dtype, void* stream);
/* plain all-reduce(sum) of a [count] T buffer, in place (AllReduceSum). */
B200_API int b200_comm_allreduce_sum(b200_comm* comm, void* data, size_t count, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decode-step driver.  Stands where LanguageModel::Run(kForward) (models/language_model.cc:493-542) +
 * UnifiedDecoder::Forward (models/llama/unified_decoder.cc:163-380) + Generation::Forward (greedy) stand:
 * one call enqueues embedding -> L x [qkv GEMM, decode attention, wo GEMM, (all-reduce+)residual RMSNorm,
 * w1w3 GEMM + gated SiLU, w2 GEMM, (all-reduce+)residual RMSNorm] -> lm_head GEMM -> argmax on `stream`
 * with no host synchronisation (CUDA-graph capturable).  Weights are caller-owned device memory in the
 * `prepared` layout; per-rank shapes when tp_size > 1 (heads, inter_size and vocab already divided).
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200_model b200_model;

typedef struct {
    int   num_layers;
    int   hidden;
    int   num_heads;    /* local q heads */
    int   num_kv_heads; /* local kv heads */
    int   head_dim;
    int   inter_size;   /* local */
    int   vocab;        /* local lm_head rows (multiple of 128) */
    int   embedding_rows; /* rows of the (replicated) embedding table */
    float norm_eps;
    float inv_sqrt_dh;
    b200_rope_config rope;
    int   dtype;
    int   quant_policy; /* KV cache: 0 / 4 / 8 */
    int   block_len;
    int   max_batch;
    int   tp_size;
    int   tp_rank;
} b200_model_config;

enum { B200_W_QKV = 0, B200_W_O = 1, B200_W_GATE_UP = 2, B200_W_DOWN = 3, B200_W_LM_HEAD = 4 };
/* B200_T_QKV_BIAS: optional per-layer [(num_heads + 2 num_kv_heads) * head_dim] T, laid out like a w_qkv output row.
 * Added to q / k / v by the attention prologue in T, where the reference adds it (AttentionParams::q_bias/k_bias/v_bias
 * from w_qkv->bias, unified_attention_layer.cc:495-497) -- Qwen2-family checkpoints. */
/* B200_T_Q_NORM / B200_T_K_NORM: optional per-layer [head_dim] T weights of the per-head q / k RMSNorm (Qwen3 family);
 * when both are set the step runs b200_qk_rmsnorm on the QKV GEMM output (eps = norm_eps). */
enum {
    B200_T_EMBEDDING  = 0,
    B200_T_FINAL_NORM = 1,
    B200_T_ATTN_NORM  = 2,
    B200_T_FFN_NORM   = 3,
    B200_T_QKV_BIAS   = 4,
    B200_T_Q_NORM     = 5,
    B200_T_K_NORM     = 6
};

typedef struct {
    int            batch_size;
    const int*     token_ids;     /* [batch] input token of every sequence */
    const int*     cu_q_len;      /* [batch+1] = 0..batch */
    const int*     cu_k_len;      /* [batch+1] context lengths INCLUDING the new token */
    const uint8_t* finished;      /* [batch] or NULL */
    void**         block_ptrs;
    const int*     cu_block_nums; /* [batch+1] */
    int            layer_stride_bytes; /* bytes of one layer inside a KV block */
    int            max_k_len;
    int*           out_tokens;    /* [batch] greedy token ids (may be NULL: logits only) */
    int            advance;       /* != 0: after sampling, token_ids <- out_tokens and k_len += 1 on device
                                     (writes through token_ids / cu_k_len; lets a captured graph self-iterate) */
} b200_step_inputs;

B200_API int b200_model_create(b200_model** model, const b200_model_config* config);
B200_API void b200_model_destroy(b200_model* model);
B200_API int b200_model_set_linear(b200_model* model, int layer, int which, const b200_linear_desc* desc);
B200_API int b200_model_set_tensor(b200_model* model, int layer, int which, const void* ptr);
B200_API int b200_model_set_comm(b200_model* model, b200_comm* comm);
/* internal activations for tests: 0 logits [max_batch, vocab], 1 residual, 2 last GEMM output, 3 qkv */
B200_API void* b200_model_buffer(b200_model* model, int which);
B200_API int b200_model_decode_step(b200_model* model, const b200_step_inputs* inputs, void* stream);
/* Mixed step (chunked prefill, "dynamic split-fuse": unified_attention_layer.cc:310, 613-651): `token_num` rows,
 * sequence b owns rows [cu_q_len[b], cu_q_len[b+1]) (its q_len new tokens, 1 for a decoding sequence);
 * inputs->token_ids and out_tokens have token_num entries (the caller samples from the LAST row of each sequence);
 * cu_k_len counts the new tokens; batch_size <= token_num <= max_batch; inputs->advance must be 0.  Attention runs
 * through b200_prefill_attention, every other operator is row-wise and shared with b200_model_decode_step. */
B200_API int b200_model_chunk_step(b200_model* model, const b200_step_inputs* inputs, int token_num, void* stream);
B200_API int b200_step_advance(const int* token_ids, const int* out_tokens, int* cu_k_len, const uint8_t* finished,
                      int batch, void* stream);

/* CUDA-IPC helpers for the one-process-per-GPU communicator (host side exchanges the 64-byte handles). */
B200_API int b200_ipc_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
B200_API int b200_ipc_open(const unsigned char handle[64], void** ptr);
B200_API int b200_ipc_close(void* ptr);
B200_API int b200_ipc_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* B200_TURBOMIND_H */
