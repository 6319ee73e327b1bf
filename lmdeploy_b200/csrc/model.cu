// Decode-step driver: the layer loop of one decode forward, natively in C++ (the reference's is C++ too:
//   LanguageModel::Impl::Forward      src/turbomind/models/language_model.cc:493-542
//   UnifiedDecoder::Forward           src/turbomind/models/llama/unified_decoder.cc:163-380
//   UnifiedAttentionLayer::Forward    src/turbomind/models/llama/unified_attention_layer.cc:365-441
//   LlamaFfnLayer::forward            src/turbomind/models/llama/LlamaFfnLayer.cc:28-91
//   Generation::Forward (greedy)      src/turbomind/generation/generation.cc:247-293 )
// One call enqueues the whole step on a stream with no host synchronisation, so the caller can capture it in a
// CUDA graph.  Activations between kernels live in the UMMA K-major tile layout so that every GEMM's B operand
// is one contiguous TMA bulk copy.
#include <cstdlib>
#include "../../include/b200_turbomind.h"
#include "host_utils.h"

#include <cstring>
#include <vector>

namespace b200 {

struct LayerW {
    b200_linear_desc qkv{}, wo{}, w13{}, w2{};
    const void*      attn_norm{};
    const void*      ffn_norm{};
    const void*      qkv_bias{};  // optional (Qwen2 family)
    const void*      q_norm{};    // optional per-head q / k RMSNorm weights (Qwen3 family)
    const void*      k_norm{};
};

}  // namespace b200

struct b200_model {
    b200_model_config       cfg{};
    std::vector<b200::LayerW> layers;
    b200_linear_desc        lm_head{};
    const void*             embedding{};
    const void*             final_norm{};
    b200_comm*              comm{};
    // device buffers
    int   mt_cap{};
    void* residual{};
    void* xp{};
    void* qkv{};
    void* attn_xp{};
    void* o{};
    void* o_x[2]{};  // GEMM outputs inside the two halves of the comm exchange buffer (TP) -- override `o`
    void* act_xp{};
    void* logits{};
    void* attn_ws{};
    size_t attn_ws_bytes{};
    void* lin_ws{};
    size_t lin_ws_bytes{};
    int*  plan{};
    float* rope_table{};
    // B200_STEP_TRACE=1: CUDA-event time per operator of every (un-captured) step, printed to stderr -- the analogue of
    // the reference's NVTX ranges around the same calls (utils/nvtx_utils.h:37-50)
    bool trace_on{};
    bool fuse_norm{false};  // B200_FUSED_NORM=1: residual+RMSNorm inside the wo / w2 GEMM epilogue (measured: no faster
                            // than the PDL-overlapped separate kernel -- 4.275 vs 4.260 ms/step -- so off by default)
    std::vector<std::tuple<int, cudaEvent_t, cudaEvent_t>> trace;
};

using namespace b200;

static int round16(int x)
{
    return std::max(16, (x + 15) / 16 * 16);
}

extern "C" int b200_model_create(b200_model** out, const b200_model_config* c)
{
    if (!out || !c) {
        return fail(B200_ERR_INVALID_ARG, "model_create: NULL argument");
    }
    if (c->head_dim != 128) {
        return fail(B200_ERR_UNSUPPORTED, "head_dim must be 128");
    }
    if (c->max_batch <= 0 || c->max_batch > 128) {
        return fail(B200_ERR_UNSUPPORTED, "max_batch must be in [1, 128]");
    }
    if (c->num_heads % c->num_kv_heads || c->hidden % 128 || c->inter_size % 128 || c->vocab % 128) {
        return fail(B200_ERR_UNSUPPORTED, "hidden / inter_size / vocab (per rank) must be multiples of 128");
    }
    auto* m   = new b200_model();
    m->trace_on = getenv("B200_STEP_TRACE") != nullptr;
    m->fuse_norm = getenv("B200_FUSED_NORM") != nullptr;
    m->cfg    = *c;
    m->layers.resize(c->num_layers);
    m->mt_cap = round16(c->max_batch);
    const size_t es = 2;
    const int    MT = m->mt_cap, MB = c->max_batch;
    const int    qkv_dim = (c->num_heads + 2 * c->num_kv_heads) * c->head_dim;
    auto alloc = [&](void** p, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) {
            return fail_cuda(e, "cudaMalloc(model buffers)");
        }
        e = cudaMemset(*p, 0, bytes);
        return e == cudaSuccess ? 0 : fail_cuda(e, "cudaMemset");
    };
    int rc = 0;
    rc = rc ? rc : alloc(&m->residual, (size_t)MB * c->hidden * es);
    rc = rc ? rc : alloc(&m->xp, (size_t)MT * c->hidden * es);
    rc = rc ? rc : alloc(&m->qkv, (size_t)MB * qkv_dim * es);
    rc = rc ? rc : alloc(&m->attn_xp, (size_t)MT * c->num_heads * c->head_dim * es);
    // `o` also receives the row-major prefill-attention output of a chunk step ([rows, num_heads * head_dim]), which
    // is wider than `hidden` for Qwen3-style geometries
    rc = rc ? rc : alloc(&m->o, (size_t)MB * std::max(c->hidden, c->num_heads * c->head_dim) * es);
    rc = rc ? rc : alloc(&m->act_xp, (size_t)MT * c->inter_size * es);
    rc = rc ? rc : alloc(&m->logits, (size_t)MB * c->vocab * es);
    m->attn_ws_bytes = b200_decode_attention_workspace_bytes(MB, c->num_heads, c->num_kv_heads);
    rc = rc ? rc : alloc(&m->attn_ws, m->attn_ws_bytes);
    // linear workspace: sized for the largest (K, N) on the path
    size_t lin = 0;
    {
        const int shapes[5][2] = {{c->hidden, qkv_dim},
                                  {c->num_heads * c->head_dim, c->hidden},
                                  {c->hidden, 2 * c->inter_size},
                                  {c->inter_size, c->hidden},
                                  {c->hidden, c->vocab}};
        for (auto& s : shapes) {
            b200_linear_desc d{};
            d.input_dim = s[0], d.output_dim = (s[1] + 127) / 128 * 128, d.weight_format = 0, d.dtype = c->dtype,
            d.group_size = 128;
            lin          = std::max(lin, b200_linear_workspace_bytes(&d, MB));
        }
    }
    m->lin_ws_bytes = lin;
    rc = rc ? rc : alloc(&m->lin_ws, lin);
    rc = rc ? rc : alloc((void**)&m->plan, (size_t)(MB + 1) * 4);
    rc = rc ? rc : alloc((void**)&m->rope_table, (size_t)MB * 64 * 8);
    if (rc) {
        b200_model_destroy(m);
        return rc;
    }
    *out = m;
    return 0;
}

extern "C" void b200_model_destroy(b200_model* m)
{
    if (!m) {
        return;
    }
    void* bufs[] = {m->residual, m->xp, m->qkv, m->attn_xp, m->o, m->act_xp, m->logits, m->attn_ws, m->lin_ws, m->plan, m->rope_table};
    for (void* p : bufs) {
        if (p) {
            cudaFree(p);
        }
    }
    delete m;
}

extern "C" int b200_model_set_linear(b200_model* m, int layer, int which, const b200_linear_desc* d)
{
    if (!m || !d) {
        return fail(B200_ERR_INVALID_ARG, "set_linear: NULL argument");
    }
    if (which == B200_W_LM_HEAD) {
        m->lm_head = *d;
        return 0;
    }
    if (layer < 0 || layer >= (int)m->layers.size()) {
        return fail(B200_ERR_INVALID_ARG, "set_linear: layer out of range");
    }
    LayerW& L = m->layers[layer];
    switch (which) {
        case B200_W_QKV: L.qkv = *d; break;
        case B200_W_O: L.wo = *d; break;
        case B200_W_GATE_UP: L.w13 = *d; break;
        case B200_W_DOWN: L.w2 = *d; break;
        default: return fail(B200_ERR_INVALID_ARG, "set_linear: bad weight id");
    }
    return 0;
}

extern "C" int b200_model_set_tensor(b200_model* m, int layer, int which, const void* ptr)
{
    if (!m || !ptr) {
        return fail(B200_ERR_INVALID_ARG, "set_tensor: NULL argument");
    }
    if (which == B200_T_EMBEDDING) {
        m->embedding = ptr;
        return 0;
    }
    if (which == B200_T_FINAL_NORM) {
        m->final_norm = ptr;
        return 0;
    }
    if (layer < 0 || layer >= (int)m->layers.size()) {
        return fail(B200_ERR_INVALID_ARG, "set_tensor: layer out of range");
    }
    if (which == B200_T_ATTN_NORM) {
        m->layers[layer].attn_norm = ptr;
    }
    else if (which == B200_T_FFN_NORM) {
        m->layers[layer].ffn_norm = ptr;
    }
    else if (which == B200_T_QKV_BIAS) {
        m->layers[layer].qkv_bias = ptr;
    }
    else if (which == B200_T_Q_NORM) {
        m->layers[layer].q_norm = ptr;
    }
    else if (which == B200_T_K_NORM) {
        m->layers[layer].k_norm = ptr;
    }
    else {
        return fail(B200_ERR_INVALID_ARG, "set_tensor: bad tensor id");
    }
    return 0;
}

extern "C" int b200_model_set_comm(b200_model* m, b200_comm* comm)
{
    if (!m) {
        return fail(B200_ERR_INVALID_ARG, "set_comm: NULL model");
    }
    m->comm = comm;
    if (comm) {
        if (int e = b200_comm_set_geometry(comm, m->cfg.max_batch, m->cfg.hidden)) {
            return e;
        }
    }
    m->o_x[0] = comm ? b200_comm_local_buffer(comm, 0) : nullptr;
    m->o_x[1] = comm ? b200_comm_local_buffer(comm, 1) : nullptr;
    return 0;
}

extern "C" void* b200_model_buffer(b200_model* m, int which)
{
    if (!m) {
        return nullptr;
    }
    switch (which) {
        case 0: return m->logits;
        case 1: return m->residual;
        case 2: return m->o;
        case 3: return m->qkv;
        default: return nullptr;
    }
}

#define TRY(x)                                                                                                         \
    do {                                                                                                               \
        int _rc = (x);                                                                                                 \
        if (_rc) {                                                                                                     \
            return _rc;                                                                                                \
        }                                                                                                              \
    } while (0)

#define TRY_T(cat, expr)                                                                                               \
    do {                                                                                                               \
        cudaEvent_t e0_{}, e1_{};                                                                                      \
        if (m->trace_on) {                                                                                             \
            cudaEventCreate(&e0_);                                                                                     \
            cudaEventCreate(&e1_);                                                                                     \
            cudaEventRecord(e0_, (cudaStream_t)stream);                                                                \
        }                                                                                                              \
        TRY(expr);                                                                                                     \
        if (m->trace_on) {                                                                                             \
            cudaEventRecord(e1_, (cudaStream_t)stream);                                                                \
            m->trace.emplace_back(cat, e0_, e1_);                                                                      \
        }                                                                                                              \
    } while (0)

static void report_trace(b200_model* m, void* stream)
{
    static const char* names[] = {"embed+norm+plan", "qkv gemm", "attention", "wo gemm", "norm/allreduce 1", "w13 gemm",
                                  "w2 gemm", "norm/allreduce 2", "lm_head", "sample+advance"};
    cudaStreamSynchronize((cudaStream_t)stream);
    double tot[10] = {0};
    for (auto& t : m->trace) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t));
        tot[std::get<0>(t)] += ms;
        cudaEventDestroy(std::get<1>(t));
        cudaEventDestroy(std::get<2>(t));
    }
    m->trace.clear();
    double sum = 0;
    for (double v : tot) sum += v;
    fprintf(stderr, "[b200 step trace rank %d] total %.3f ms |", m->cfg.tp_rank, sum);
    for (int i = 0; i < 10; ++i) fprintf(stderr, " %s %.3f |", names[i], tot[i]);
    fprintf(stderr, "\n");
}

// One forward step over R token rows of B sequences.  Decode: R == B, one row per sequence.  Chunked prefill /
// mixed step (`chunk`): sequence b owns rows [cu_q_len[b], cu_q_len[b+1]); the attention of the step runs through
// b200_prefill_attention (causal over each sequence's pages), everything else is row-wise and identical.
static int run_step(b200_model* m, const b200_step_inputs* in, int R, bool chunk, void* stream)
{
    const b200_model_config& c = m->cfg;
    const int                B = in->batch_size;
    if (!m->embedding || !m->final_norm || !m->lm_head.prepared) {
        return fail(B200_ERR_STATE, "decode_step: model weights not set");
    }
    const int MT      = round16(R);
    const int qkv_dim = (c.num_heads + 2 * c.num_kv_heads) * c.head_dim;
    const int es      = 2;
    const int L       = c.num_layers;
    void*     o_buf   = m->o_x[0] ? m->o_x[0] : m->o;   // wo output (exchange half 0 under TP)
    void*     o_buf2  = m->o_x[1] ? m->o_x[1] : m->o;   // w2 output (exchange half 1 under TP)

    TRY_T(0, b200_embedding_lookup(m->residual, m->embedding, in->token_ids, R, c.hidden, c.embedding_rows, c.dtype, stream));
    TRY_T(0, b200_rmsnorm(m->xp, MT, m->residual, m->layers[0].attn_norm, R, c.hidden, c.norm_eps, c.dtype, stream));
    if (!chunk) {
        TRY_T(0, b200_decode_attention_plan(m->plan, in->cu_k_len, in->finished, B, stream));
    }
    TRY_T(0, b200_rope_table(m->rope_table, &c.rope, in->cu_q_len, in->cu_k_len, B, R, stream));

    for (int l = 0; l < L; ++l) {
        const LayerW& W = m->layers[l];
        if (!W.qkv.prepared || !W.wo.prepared || !W.w13.prepared || !W.w2.prepared || !W.attn_norm || !W.ffn_norm) {
            return fail(B200_ERR_STATE, "decode_step: layer weights not set");
        }
        // ---- attention block ----
        TRY_T(1, b200_linear_forward_packed(&W.qkv, m->xp, MT, R, m->qkv, qkv_dim, 0, 0, m->lin_ws, m->lin_ws_bytes, stream));
        if (W.q_norm && W.k_norm) {
            TRY_T(1, b200_qk_rmsnorm(m->qkv, qkv_dim, W.q_norm, W.k_norm, R, c.num_heads, c.num_kv_heads, c.head_dim, c.norm_eps,
                                     c.dtype, stream));
        }
        b200_attention_params a{};
        a.out    = m->attn_xp;
        a.q      = m->qkv;
        a.k      = (char*)m->qkv + (size_t)c.num_heads * c.head_dim * es;
        a.v      = (char*)m->qkv + (size_t)(c.num_heads + c.num_kv_heads) * c.head_dim * es;
        a.stride = qkv_dim;
        if (W.qkv_bias) {
            a.q_bias = W.qkv_bias;
            a.k_bias = (const char*)W.qkv_bias + (size_t)c.num_heads * c.head_dim * es;
            a.v_bias = (const char*)W.qkv_bias + (size_t)(c.num_heads + c.num_kv_heads) * c.head_dim * es;
        }
        a.cu_q_len = in->cu_q_len, a.cu_k_len = in->cu_k_len, a.finished = in->finished;
        a.block_ptrs = in->block_ptrs, a.cu_block_nums = in->cu_block_nums;
        a.layer_offset = l * in->layer_stride_bytes, a.block_len = c.block_len;
        a.token_num = R, a.batch_size = B, a.max_k_len = in->max_k_len;
        a.num_heads = c.num_heads, a.num_kv_heads = c.num_kv_heads, a.size_per_head = c.head_dim;
        a.inv_sqrt_dh = c.inv_sqrt_dh, a.window_size = 0;
        a.rope = c.rope, a.quant_policy = c.quant_policy, a.dtype = c.dtype;
        a.workspace = m->attn_ws, a.workspace_bytes = m->attn_ws_bytes;
        a.plan = m->plan, a.rope_table = m->rope_table, a.out_packed_mt = MT, a.stream = stream;
        if (!chunk) {
            TRY_T(2, b200_decode_attention(&a));
        }
        else {
            // row-major into the (currently idle) GEMM output buffer, then into the packed layout the wo GEMM reads
            const int attn_dim = c.num_heads * c.head_dim;
            a.out = m->o, a.out_packed_mt = 0, a.plan = nullptr;
            TRY_T(2, b200_prefill_attention(&a));
            TRY_T(2, b200_pack_activations(m->o, attn_dim, R, attn_dim, MT, c.dtype, m->attn_xp, stream));
        }
        // wo GEMM + AllreduceResidualRMSnorm #1 (unified_decoder.cc:278-287).  Single GPU: the residual add and the
        // RMSNorm are fused into the GEMM's split-K epilogue when the launch plan has one.
        bool fused1 = false;
        if (!m->comm && m->fuse_norm) {
            const int rc = b200_linear_forward_packed_residual_rmsnorm(&W.wo, m->attn_xp, MT, R, m->residual, W.ffn_norm,
                                                                       c.norm_eps, m->xp, MT, m->lin_ws, m->lin_ws_bytes,
                                                                       stream);
            if (rc == 0) {
                fused1 = true;
            }
            else if (rc != B200_ERR_UNSUPPORTED) {
                return rc;
            }
        }
        if (!fused1) {
            TRY_T(3, b200_linear_forward_packed(&W.wo, m->attn_xp, MT, R, o_buf, c.hidden, 0, 0, m->lin_ws, m->lin_ws_bytes, stream));
            if (m->comm) {
                TRY_T(4, b200_comm_allreduce_residual_rmsnorm(m->comm, m->xp, MT, m->residual, o_buf, W.ffn_norm, nullptr, R,
                                                              c.hidden, c.norm_eps, c.dtype, stream));
            }
            else {
                TRY_T(4, b200_residual_rmsnorm(m->xp, MT, m->residual, o_buf, W.ffn_norm, nullptr, R, c.hidden, c.norm_eps,
                                               c.dtype, stream));
            }
        }
        // ---- FFN block ----
        TRY_T(5, b200_linear_forward_packed(&W.w13, m->xp, MT, R, m->act_xp, c.inter_size, 1, MT, m->lin_ws, m->lin_ws_bytes,
                                       stream));
        const void* next_norm = (l + 1 < L) ? m->layers[l + 1].attn_norm : m->final_norm;
        bool        fused2    = false;
        if (!m->comm && m->fuse_norm) {
            const int rc = b200_linear_forward_packed_residual_rmsnorm(&W.w2, m->act_xp, MT, R, m->residual, next_norm,
                                                                       c.norm_eps, m->xp, MT, m->lin_ws, m->lin_ws_bytes,
                                                                       stream);
            if (rc == 0) {
                fused2 = true;
            }
            else if (rc != B200_ERR_UNSUPPORTED) {
                return rc;
            }
        }
        if (!fused2) {
            TRY_T(6, b200_linear_forward_packed(&W.w2, m->act_xp, MT, R, o_buf2, c.hidden, 0, 0, m->lin_ws, m->lin_ws_bytes, stream));
            if (m->comm) {
                TRY_T(7, b200_comm_allreduce_residual_rmsnorm(m->comm, m->xp, MT, m->residual, o_buf2, next_norm, nullptr, R,
                                                              c.hidden, c.norm_eps, c.dtype, stream));
            }
            else {
                TRY_T(7, b200_residual_rmsnorm(m->xp, MT, m->residual, o_buf2, next_norm, nullptr, R, c.hidden, c.norm_eps, c.dtype,
                                               stream));
            }
        }
    }
    // PostEmbedding (language_model.cc:285-335) + greedy sampling
    TRY_T(8, b200_linear_forward_packed(&m->lm_head, m->xp, MT, R, m->logits, c.vocab, 0, 0, m->lin_ws, m->lin_ws_bytes,
                                   stream));
    if (in->out_tokens) {
        // c.vocab is the per-rank lm_head width padded to a multiple of 128; only the columns below the real vocabulary
        // size (= rows of the embedding table) may be sampled (sampling.cc:132: vocab_size, not vocab_size_padded)
        const int valid = std::max(0, std::min(c.vocab, c.embedding_rows - c.tp_rank * c.vocab));
        if (m->comm && c.tp_size > 1) {
            TRY_T(9, b200_comm_argmax_allgather(m->comm, in->out_tokens, m->logits, R, c.vocab, valid, c.dtype, stream));
        }
        else {
            TRY_T(9, b200_argmax(in->out_tokens, m->logits, R, valid, c.vocab, c.dtype, stream));
        }
    }
    if (in->advance && !chunk) {
        TRY_T(9, b200_step_advance(in->token_ids, in->out_tokens, (int*)in->cu_k_len, in->finished, B, stream));
    }
    if (m->trace_on) {
        report_trace(m, stream);
    }
    return 0;
}

extern "C" int b200_model_decode_step(b200_model* m, const b200_step_inputs* in, void* stream)
{
    if (!m || !in) {
        return fail(B200_ERR_INVALID_ARG, "decode_step: NULL argument");
    }
    if (in->batch_size <= 0 || in->batch_size > m->cfg.max_batch) {
        return fail(B200_ERR_INVALID_ARG, "decode_step: batch_size out of range");
    }
    return run_step(m, in, in->batch_size, false, stream);
}

extern "C" int b200_model_chunk_step(b200_model* m, const b200_step_inputs* in, int token_num, void* stream)
{
    if (!m || !in) {
        return fail(B200_ERR_INVALID_ARG, "chunk_step: NULL argument");
    }
    if (in->batch_size <= 0 || token_num < in->batch_size || token_num > m->cfg.max_batch) {
        return fail(B200_ERR_INVALID_ARG, "chunk_step: need batch_size <= token_num <= max_batch rows");
    }
    if (in->advance) {
        return fail(B200_ERR_UNSUPPORTED, "chunk_step: device-side advance is a decode-only feature");
    }
    return run_step(m, in, token_num, true, stream);
}
