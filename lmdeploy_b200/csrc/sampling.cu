// Token sampling beyond greedy: repetition penalty -> temperature -> top-k -> softmax -> top-p / min-p -> draw.
//
// Replaces, behind the C ABI, the reference's generation stage for do_sample requests
//   LogitsProcessor::Forward      src/turbomind/generation/logits_processor.cc:66-125
//     RepetitionPenaltyKernel       kernels/sampling_penalty_kernels.cu:135-170   (logit < 0 ? logit * p : logit / p)
//     batchApplyTemperaturePenalty  kernels/sampling_penalty_kernels.cu:27-80     (logit * fdividef(1, T + 1e-6))
//   Sampling::Forward             src/turbomind/generation/sampling.cc:92-183
//     topKSortStage1/2              kernels/sampling_topk_kernels.cu:97-215  (k largest, descending; p = exp(l - max) / sum)
//     topPMinPFilter                kernels/sampling_topp_kernels.cu:290-383 (prefix until cum > top_p; p < p_max * min_p out)
//     sampling                      kernels/sampling_kernels.cu:15-57        (curand_uniform; first prefix sum > u)
//   InitializeRandomStates        kernels/sampling_topk_kernels.cu:74-95   (curand_init(seed, 0, 0): XORWOW)
// It is not a port: ONE kernel per step, one CTA per sampled row.  The k largest logits are found with an 8-bit radix select
// over order-preserving keys (4 passes over the row, no k-pass arg-max loop), gathered in vocabulary order and sorted in
// shared memory (bitonic, value descending / index ascending), after which the reference's arithmetic is followed on the
// sorted candidates: exp(l - max) summed sequentially in sorted order, top-p cut at the first prefix sum > top_p, min-p
// against p_max, renormalisation, curand_uniform from the sequence's own XORWOW state, first prefix sum > u.
// top_k = 0 (no top-k filter; rare: lmdeploy's default is 50) keeps the whole vocabulary: the top-p cut is then found by
// bisection on the probability value (31 deterministic block reductions) and the draw runs in vocabulary order -- the
// same candidate set and the same distribution as sorting 92k entries, without the sort.
#include "../../include/b200_turbomind.h"
#include "common.cuh"
#include "host_utils.h"

#include <curand_kernel.h>

namespace b200 {
namespace sampling {

constexpr int kThreads = 1024;
constexpr int kMaxTopK = 1024;  // the reference's limit (sampling_topk_kernels.cu:262-264)

struct Params {
    const void*         logits;
    int64_t             ld;
    int                 vocab;
    const int*          rows;
    const float*        temperature;
    const int*          top_k;
    const float*        top_p;
    const float*        min_p;
    const float*        repetition_penalty;
    const int* const*   history;
    const int*          history_len;
    curandState_t*      rng;
    const int*          rng_index;
    int*                out_ids;
    float*              out_uniform;
    int*                out_kept;
    float*              ws;  // [count][vocab_pad]
    int                 vocab_pad;
};

// order-preserving float -> uint key (larger float <=> larger key); -inf / NaN-free inputs assumed finite or -inf
B200_DEVICE uint32_t key_of(float x)
{
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

B200_DEVICE float block_max(float v, float* red)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < kThreads / 32; ++w) r = fmaxf(r, red[w]);  // fixed order
    return r;
}

B200_DEVICE float block_sum(float v, float* red)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) r += red[w];  // fixed order: deterministic
    return r;
}

// exclusive prefix of `v` over the block in thread order, plus the block total
B200_DEVICE int block_excl_scan(int v, int* red, int& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int       inc  = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    __syncthreads();
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) base += red[w];
        tot += red[w];
    }
    total = tot;
    return base + inc - v;
}

B200_DEVICE float block_excl_scan_f(float v, float* red, float& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float     inc  = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    __syncthreads();
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    float base = 0.f, tot = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) base += red[w];
        tot += red[w];
    }
    total = tot;
    return base + inc - v;
}

template<class T>
__global__ void __launch_bounds__(kThreads) sample_kernel(const Params p)
{
    using Ops = TypeOps<T>;
    extern __shared__ uint32_t mask[];  // repetition-penalty bitmask over the vocabulary (vocab / 32 words)
    __shared__ float    s_val[kMaxTopK];
    __shared__ int      s_idx[kMaxTopK];
    __shared__ float    s_redf[kThreads / 32];
    __shared__ int      s_redi[kThreads / 32];
    __shared__ int      s_hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int      s_need;
    __shared__ float    s_u, s_sum, s_thr;
    __shared__ int      s_kept, s_pick;

    const int    item  = blockIdx.x;
    const int    tid   = threadIdx.x;
    const int    V     = p.vocab;
    const int    row   = p.rows ? p.rows[item] : item;
    const T*     lg    = reinterpret_cast<const T*>(p.logits) + (size_t)row * p.ld;
    float*       x     = p.ws + (size_t)item * p.vocab_pad;
    const float  top_p = p.top_p ? p.top_p[item] : 1.f;
    const float  min_p = p.min_p ? p.min_p[item] : 0.f;
    int          k     = p.top_k ? p.top_k[item] : 0;
    k                  = (k <= 0 || k > V) ? 0 : min(k, kMaxTopK);

    // ---- random number first (thread 0; the state advances exactly once per sampled token, as in the reference) ----
    if (tid == 0) {
        curandState_t* st = p.rng + (p.rng_index ? p.rng_index[item] : item);
        s_u               = curand_uniform(st);
        if (p.out_uniform) p.out_uniform[item] = s_u;
    }

    // ---- repetition penalty mask ----
    const float pen     = p.repetition_penalty ? p.repetition_penalty[item] : 1.f;
    const bool  use_pen = pen != 1.f && p.history && p.history[item] && p.history_len[item] > 0;
    if (use_pen) {
        const int words = (V + 31) / 32;
        for (int i = tid; i < words; i += kThreads) mask[i] = 0;
        __syncthreads();
        const int* h = p.history[item];
        for (int i = tid; i < p.history_len[item]; i += kThreads) {
            const int t = h[i];
            if (t >= 0 && t < V) atomicOr(&mask[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
    }

    // ---- logits -> processed fp32 row, running maximum ----
    const float scale = __fdividef(1.f, (p.temperature ? p.temperature[item] : 1.f) + 1e-6f);
    float       mx    = -INFINITY;
    for (int i = tid; i < V; i += kThreads) {
        float v = Ops::to_f32(lg[i]);
        if (use_pen && (mask[i >> 5] >> (i & 31) & 1u)) {
            v = v < 0.f ? v * pen : v / pen;
        }
        v *= scale;
        x[i] = v;
        mx   = fmaxf(mx, v);
    }
    mx = block_max(mx, s_redf);

    if (k > 0) {
        // ================= top-k: radix select of the k-th largest key, gather, sort, the reference's arithmetic =================
        uint32_t prefix = 0;  // bits fixed so far (from the top)
        int      need   = k;  // rank still to locate inside the current prefix bucket
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kThreads) s_hist[i] = 0;
            __syncthreads();
            const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = tid; i < V; i += kThreads) {
                const uint32_t key = key_of(x[i]);
                if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0, b = 255;
                for (; b > 0; --b) {
                    if (acc + s_hist[b] >= need) break;
                    acc += s_hist[b];
                }
                s_prefix = prefix | ((uint32_t)b << shift);
                s_need   = need - acc;
            }
            __syncthreads();
            prefix = s_prefix;
            need   = s_need;
        }
        // prefix = key of the k-th largest value; `need` = how many entries EQUAL to it are kept (lowest indices first)
        // gather the kept entries in vocabulary order: chunks of kThreads consecutive indices, one exclusive scan each
        int base_gt = 0, base_eq = 0;  // running totals over the chunks already visited
        for (int c0 = 0; c0 < V; c0 += kThreads) {
            const int      i   = c0 + tid;
            const uint32_t key = i < V ? key_of(x[i]) : 0u;
            const int      gt  = (i < V && key > prefix) ? 1 : 0;
            const int      eq  = (i < V && key == prefix) ? 1 : 0;
            int            tot_gt, tot_eq;
            const int      r_gt = block_excl_scan(gt, s_redi, tot_gt);
            const int      r_eq = block_excl_scan(eq, s_redi, tot_eq);
            const int      eq_rank = base_eq + r_eq;  // rank of this entry among the equal ones
            // slot = (# kept greater before me) + (# kept equal before me)
            if (gt || (eq && eq_rank < need)) {
                const int slot = base_gt + r_gt + min(eq_rank, need);
                s_val[slot]    = x[i];
                s_idx[slot]    = i;
            }
            base_gt += tot_gt;
            base_eq += tot_eq;
            __syncthreads();
        }
        // pad to a power of two and sort: value descending, index ascending on equal values (deterministic)
        int n2 = 1;
        while (n2 < k) n2 <<= 1;
        for (int i = k + tid; i < n2; i += kThreads) {
            s_val[i] = -INFINITY;
            s_idx[i] = 0x7fffffff;
        }
        __syncthreads();
        for (int size = 2; size <= n2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < n2 / 2; t += kThreads) {
                    const int  lo  = 2 * t - (t & (stride - 1));  // index with bit `stride` clear
                    const int  hi  = lo + stride;
                    const bool asc = (lo & size) != 0;  // sub-sequence direction (overall: descending)
                    const float va = s_val[lo], vb = s_val[hi];
                    const int   ia = s_idx[lo], ib = s_idx[hi];
                    const bool  a_first = va > vb || (va == vb && ia < ib);  // "a" belongs before "b" in the final order
                    if (a_first == asc) {
                        s_val[lo] = vb, s_val[hi] = va;
                        s_idx[lo] = ib, s_idx[hi] = ia;
                    }
                }
                __syncthreads();
            }
        }
        // p_i = exp(l_i - l_0), summed in sorted order by one thread (topKSortStage2), then top-p / min-p / draw on the list
        if (tid == 0) {
            const float l0  = s_val[0];
            float       sum = 0.f;
            for (int i = 0; i < k; ++i) {
                const float e = __expf(s_val[i] - l0);
                s_val[i]      = e;
                sum += e;
            }
            for (int i = 0; i < k; ++i) s_val[i] = s_val[i] / sum;
            int   kept  = k;
            float total = 1.f;
            if (top_p != 1.f) {  // first prefix sum > top_p closes the set
                float acc = 0.f;
                kept      = k;
                total     = 0.f;
                for (int i = 0; i < k; ++i) {
                    acc += s_val[i];
                    if (acc > top_p || i == k - 1) {
                        kept  = i + 1;
                        total = acc;
                        break;
                    }
                }
            }
            if (min_p != 0.f) {  // entries below p_max * min_p (renormalised by the running total) leave
                const float kEps   = 1e-6f;
                const float scaled = s_val[0] / (total + kEps) * min_p;
                float       acc    = 0.f;
                int         i      = 0;
                for (; i < kept; ++i) {
                    const float q = s_val[i] / (total + kEps);
                    if (q < scaled) break;
                    acc += q;
                }
                if (i == 0) {  // cannot happen for min_p <= 1 (the first entry IS p_max); keep one candidate regardless
                    i   = 1;
                    acc = s_val[0] / (total + kEps);
                }
                kept = i;
                total *= acc;
            }
            // draw: first prefix sum of the renormalised probabilities > u
            const float u   = s_u;
            float       acc = 0.f;
            int         sel = kept - 1;
            for (int i = 0; i < kept; ++i) {
                acc += s_val[i] / total;
                if (acc > u) {
                    sel = i;
                    break;
                }
            }
            p.out_ids[item] = s_idx[sel];
            if (p.out_kept) p.out_kept[item] = kept;
        }
        return;
    }

    // ================= no top-k: the whole vocabulary is the candidate list =================
    float part = 0.f;
    for (int i = tid; i < V; i += kThreads) {
        const float e = __expf(x[i] - mx);
        x[i]          = e;
        part += e;
    }
    const float denom = block_sum(part, s_redf);
    for (int i = tid; i < V; i += kThreads) x[i] = x[i] / denom;  // probabilities; p_max = 1 / denom
    __syncthreads();
    const float pmax = 1.f / denom;
    // threshold t* = the largest probability value with  sum_{p_i >= t} p_i > top_p  (the value at which the sorted prefix
    // closes); bisection on the bit pattern (positive floats order like their bits)
    float thr = 0.f;
    if (top_p != 1.f) {
        uint32_t lo = 0u, hi = __float_as_uint(pmax) + 1u;  // cond(lo) holds (sum of everything = 1 > top_p), cond(hi) fails
        while (hi - lo > 1u) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float          s   = 0.f;
            for (int i = tid; i < V; i += kThreads) {
                const float q = x[i];
                s += (__float_as_uint(q) >= mid) ? q : 0.f;
            }
            s = block_sum(s, s_redf);
            if (s > top_p) {
                lo = mid;
            }
            else {
                hi = mid;
            }
        }
        thr = __uint_as_float(lo);
    }
    // entries strictly above the threshold are all in; of the ones equal to it, as many (lowest indices) as the prefix needs
    float gsum = 0.f;
    int   geq  = 0;
    for (int i = tid; i < V; i += kThreads) {
        const float q = x[i];
        gsum += q > thr ? q : 0.f;
        geq += (q == thr && thr > 0.f) ? 1 : 0;
    }
    gsum = block_sum(gsum, s_redf);
    int need_eq = 0;
    if (top_p != 1.f && thr > 0.f) {
        // smallest m with gsum + m * thr > top_p
        need_eq = (int)floorf((top_p - gsum) / thr) + 1;
        need_eq = max(1, need_eq);
    }
    const float minp_cut = min_p != 0.f ? pmax * min_p : 0.f;
    // total mass of the final candidate set, then the draw in vocabulary order
    float total    = 0.f;
    int   kept     = 0;
    int   base_eq  = 0;
    // pass 1: mass
    for (int c0 = 0; c0 < V; c0 += kThreads) {
        const int   i  = c0 + tid;
        const float q  = i < V ? x[i] : 0.f;
        const int   eq = (i < V && top_p != 1.f && q == thr && thr > 0.f) ? 1 : 0;
        int         tot_eq;
        const int   r_eq = block_excl_scan(eq, s_redi, tot_eq);
        bool        in   = i < V && (top_p == 1.f || q > thr || (eq && base_eq + r_eq < need_eq));
        in               = in && q >= minp_cut;
        float tot;
        block_excl_scan_f(in ? q : 0.f, s_redf, tot);
        int totk;
        block_excl_scan(in ? 1 : 0, s_redi, totk);
        total += tot;
        kept += totk;
        base_eq += tot_eq;
    }
    // pass 2: first prefix sum (in vocabulary order) of q / total > u
    const float u = s_u;
    if (tid == 0) s_pick = -1;
    __syncthreads();
    float run  = 0.f;
    int   last = -1;
    base_eq    = 0;
    for (int c0 = 0; c0 < V && s_pick < 0; c0 += kThreads) {
        const int   i  = c0 + tid;
        const float q  = i < V ? x[i] : 0.f;
        const int   eq = (i < V && top_p != 1.f && q == thr && thr > 0.f) ? 1 : 0;
        int         tot_eq;
        const int   r_eq = block_excl_scan(eq, s_redi, tot_eq);
        bool        in   = i < V && (top_p == 1.f || q > thr || (eq && base_eq + r_eq < need_eq));
        in               = in && q >= minp_cut;
        const float w    = in ? q / total : 0.f;
        float       tot;
        const float ex   = block_excl_scan_f(w, s_redf, tot);
        const bool  hit  = in && (run + ex + w > u);
        // the first hit in vocabulary order wins
        if (tid == 0) s_hist[0] = 0x7fffffff;
        __syncthreads();
        if (hit) atomicMin(&s_hist[0], i);
        if (in) last = i;
        __syncthreads();
        if (tid == 0 && s_hist[0] != 0x7fffffff) s_pick = s_hist[0];
        __syncthreads();
        run += tot;
        base_eq += tot_eq;
    }
    if (s_pick < 0) {
        // rounding left the last prefix sum <= u: the last candidate (sampling_kernels.cu: selected = min(i, n - 1))
        int mine = last;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, m));
        __syncthreads();
        if ((tid & 31) == 0) s_redi[tid >> 5] = mine;
        __syncthreads();
        if (tid == 0) {
            int r = -1;
            for (int w = 0; w < kThreads / 32; ++w) r = max(r, s_redi[w]);
            s_pick = r;
        }
        __syncthreads();
    }
    if (tid == 0) {
        p.out_ids[item] = s_pick;
        if (p.out_kept) p.out_kept[item] = kept;
    }
}

__global__ void init_rng_kernel(curandState_t* states, const unsigned long long* seeds, const uint8_t* mask, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (!mask || mask[i])) {
        curand_init(seeds[i], 0, 0, &states[i]);
    }
}

}  // namespace sampling
}  // namespace b200

using namespace b200;

extern "C" size_t b200_sampling_rng_state_bytes(void)
{
    return sizeof(curandState_t);
}

extern "C" size_t b200_sampling_workspace_bytes(int count, int vocab)
{
    if (count <= 0 || vocab <= 0) {
        return 0;
    }
    return (size_t)count * ((vocab + 3) & ~3) * sizeof(float);
}

extern "C" int b200_sampling_init_rng(void* states, const unsigned long long* seeds, const uint8_t* mask, int n, void* stream)
{
    if (!states || !seeds || n <= 0) {
        return fail(B200_ERR_INVALID_ARG, "sampling_init_rng: states / seeds NULL or n <= 0");
    }
    sampling::init_rng_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((curandState_t*)states, seeds, mask, n);
    return check_launch("init_rng_kernel");
}

extern "C" int b200_sample(const b200_sampling_params* a)
{
    if (!a || !a->logits || !a->out_ids || !a->rng_states || !a->workspace) {
        return fail(B200_ERR_INVALID_ARG, "sample: logits / out_ids / rng_states / workspace is NULL");
    }
    if (a->count <= 0 || a->vocab <= 0 || a->ld < a->vocab) {
        return fail(B200_ERR_INVALID_ARG, "sample: count, vocab must be positive and ld >= vocab");
    }
    if (a->dtype != B200_DTYPE_F16 && a->dtype != B200_DTYPE_BF16) {
        return fail(B200_ERR_INVALID_ARG, "sample: dtype must be 0 (f16) or 1 (bf16)");
    }
    if (a->workspace_bytes < b200_sampling_workspace_bytes(a->count, a->vocab)) {
        return fail(B200_ERR_WORKSPACE, "sample: workspace too small (see b200_sampling_workspace_bytes)");
    }
    if (a->repetition_penalty && (!a->history || !a->history_len)) {
        return fail(B200_ERR_INVALID_ARG, "sample: repetition_penalty needs history / history_len");
    }
    sampling::Params p{};
    p.logits = a->logits, p.ld = a->ld, p.vocab = a->vocab, p.rows = a->rows;
    p.temperature = a->temperature, p.top_k = a->top_k, p.top_p = a->top_p, p.min_p = a->min_p;
    p.repetition_penalty = a->repetition_penalty, p.history = a->history, p.history_len = a->history_len;
    p.rng = (curandState_t*)a->rng_states, p.rng_index = a->rng_index;
    p.out_ids = a->out_ids, p.out_uniform = a->out_uniform, p.out_kept = a->out_kept;
    p.ws = (float*)a->workspace, p.vocab_pad = (a->vocab + 3) & ~3;
    const size_t smem = a->repetition_penalty ? (size_t)((a->vocab + 31) / 32) * 4 : 0;
    if (smem > 160 * 1024) {
        return fail(B200_ERR_UNSUPPORTED, "sample: vocabulary too large for the repetition-penalty mask");
    }
    cudaStream_t st = (cudaStream_t)a->stream;
    if (a->dtype == B200_DTYPE_F16) {
        if (smem > 40 * 1024) {
            cudaFuncSetAttribute(sampling::sample_kernel<half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        sampling::sample_kernel<half><<<a->count, sampling::kThreads, smem, st>>>(p);
    }
    else {
        if (smem > 40 * 1024) {
            cudaFuncSetAttribute(sampling::sample_kernel<nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        sampling::sample_kernel<nv_bfloat16><<<a->count, sampling::kThreads, smem, st>>>(p);
    }
    return check_launch("sample_kernel");
}
