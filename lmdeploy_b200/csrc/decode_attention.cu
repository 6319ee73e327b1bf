// Host side + prologue/plan kernels of the decode attention path.  See decode_attention.cuh.
#include "../../include/b200_turbomind.h"
#include "decode_attention.cuh"
#include "host_utils.h"
#include "prefill_attention.cuh"

#include <cmath>
#include <mutex>

namespace b200 {
namespace attn {

static RopeConst make_rope_const(const b200_rope_config& r)
{
    RopeConst c{};
    c.type              = r.type;
    c.dim               = r.dim;
    c.scale_factor      = -std::log2(r.base) / r.dim;  // float log2, float divide
    c.inv_factor        = (r.type == 3) ? 1.f : ((r.factor != 0.f) ? (float)(1.0 / r.factor) : 1.f);
    c.attention_scaling = 1.f;
    c.exact             = r.exact_math;
    if (r.type == 5) {
        float inv_diff = (float)(1.0 / (r.llama3_high_freq_factor - r.llama3_low_freq_factor));
        c.llama3_alpha = (float)(r.llama3_original_max_position_embeddings / (2 * 3.14159265358979323846) * inv_diff);
        c.llama3_beta  = r.llama3_low_freq_factor * inv_diff;
    }
    else if (r.type == 4) {
        const double PI   = 3.14159265358979323846;
        auto         corr = [&](float num_rot) {
            return (r.dim * std::log(r.max_position_embeddings / (num_rot * 2 * PI))) / (2 * std::log(r.base));
        };
        float low  = std::max((float)std::floor(corr(r.yarn_beta_fast)), 0.f);
        float high = std::min((float)std::ceil(corr(r.yarn_beta_slow)), r.dim - 1.f);
        if (low == high) {
            high += 0.001f;
        }
        c.yarn_div2         = (float)(1.0 / (high - low) / 2.0);
        c.yarn_mulmin       = (float)(1.0 / (high - low) * low);
        c.attention_scaling = r.yarn_attention_factor;
    }
    return c;
}

// ------------------------------------------------------------------------------------------------
// plan: tiles per sequence prefix + (optionally) cos/sin table of the decode positions
// ------------------------------------------------------------------------------------------------
__global__ void plan_kernel(int* plan, const int* cu_k_len, const uint8_t* finished, int batch, int* zero, int nzero)
{
    pdl_launch_dependents();
    pdl_wait();
    for (int i = threadIdx.x; i < nzero; i += blockDim.x) {
        zero[i] = 0;  // arrival counters of the in-kernel fix-up (a launch that was aborted may have left some armed)
    }
    // one block; serial scan by chunks of blockDim
    __shared__ int s_scan[1024];
    __shared__ int s_base;
    if (threadIdx.x == 0) {
        s_base  = 0;
        plan[0] = 0;
    }
    __syncthreads();
    for (int b0 = 0; b0 < batch; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        int       v = 0;
        if (b < batch && !(finished && finished[b])) {
            v = (cu_k_len[b + 1] - cu_k_len[b] + kTile - 1) / kTile;
        }
        s_scan[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < blockDim.x; off <<= 1) {
            int x = (threadIdx.x >= off) ? s_scan[threadIdx.x - off] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += x;
            __syncthreads();
        }
        if (b < batch) {
            plan[b + 1] = s_base + s_scan[threadIdx.x];
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) {
            s_base += s_scan[threadIdx.x];
        }
        __syncthreads();
    }
}

// cos/sin of every new token's position, computed once per step and shared by all layers
__global__ void rope_table_kernel(float2* table, RopeConst rope, const int* cu_q_len, const int* cu_k_len, int batch)
{
    pdl_launch_dependents();
    pdl_wait();
    const int tok = blockIdx.x;
    int       lo = 0, hi = batch;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cu_q_len[mid] <= tok) {
            lo = mid;
        }
        else {
            hi = mid;
        }
    }
    const int b  = lo;
    const int ti = (cu_k_len[b + 1] - cu_k_len[b]) - (cu_q_len[b + 1] - cu_q_len[b]) + (tok - cu_q_len[b]);
    const int i  = threadIdx.x;  // pair index
    table[(size_t)tok * (kHeadDim / 2) + i] =
        (rope.type != 0 && 2 * i < rope.dim) ? rope_cos_sin(rope, 2 * i, ti) : make_float2(1.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// prologue: bias + RoPE (+ quantise + page store) for every new token
//   (attention_universal.h:168-348 for decode, kv_cache_utils_v2.cu:18-210 for prefill)
// ------------------------------------------------------------------------------------------------
struct PrologueParams {
    const void* q;
    const void* k;
    const void* v;
    int64_t     stride;
    const void* q_bias;
    const void* k_bias;
    const void* v_bias;
    void*       q_rot;  // [token][Hq][D] or null
    const int*  cu_q_len;
    const int*  cu_k_len;
    const int*  readonly_block_num;
    const uint8_t* finished;
    char**      block_ptrs;
    const int*  cu_block_nums;
    int         layer_offset;
    int         block_len;
    int         token_num;
    int         batch;
    int         num_heads;
    int         num_kv_heads;
    RopeConst   rope;
    const float2* rope_table;  // [token][kHeadDim/2] (cos, sin) or null
    int           wait_first;  // trigger the dependent grid only after griddepcontrol.wait (KernelParams::early)
};

template<class T, int KVBITS>
__global__ void __launch_bounds__(128) kv_prologue_kernel(const PrologueParams p)
{
    if (p.wait_first) {
        pdl_wait();
        pdl_launch_dependents();
    }
    else {
        pdl_launch_dependents();
        pdl_wait();
    }
    using Ops      = TypeOps<T>;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.y * 4 + warp;
    const int tok  = blockIdx.x;
    const int Hq   = p.q_rot ? p.num_heads : 0;
    if (item >= Hq + p.num_kv_heads) {
        return;
    }
    // token -> (sequence, q index)
    int lo = 0, hi = p.batch;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.cu_q_len[mid] <= tok) {
            lo = mid;
        }
        else {
            hi = mid;
        }
    }
    const int b = lo;
    if (p.finished && p.finished[b]) {
        return;
    }
    const int qi      = tok - p.cu_q_len[b];
    const int q_len   = p.cu_q_len[b + 1] - p.cu_q_len[b];
    const int k_len   = p.cu_k_len[b + 1] - p.cu_k_len[b];
    const int history = k_len - q_len;
    const int ti      = history + qi;
    const int di      = lane * 4;

    float2 cs[2];
    if (p.rope_table) {
        const float4 v = *reinterpret_cast<const float4*>(p.rope_table + (size_t)tok * (kHeadDim / 2) + (di >> 1));
        cs[0]          = make_float2(v.x, v.y);
        cs[1]          = make_float2(v.z, v.w);
    }
    else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            cs[i] = (p.rope.type != 0 && di + 2 * i < p.rope.dim) ? rope_cos_sin(p.rope, di + 2 * i, ti) :
                                                                     make_float2(1.f, 0.f);
        }
    }
    auto rotate = [&](T(&x)[4]) {
        if (p.rope.type == 0) {
            return;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (di + 2 * i < p.rope.dim) {
                const T c = Ops::from_f32(cs[i].x), s = Ops::from_f32(cs[i].y);
                const T a = x[2 * i], bb = x[2 * i + 1];
                // the reference binary contracts `c*x0 - s*x1` to HMUL2 + HFMA2 (SASS of oracle/_ref)
                x[2 * i]     = Ops::fma(c, a, Ops::neg(Ops::mul(s, bb)));
                x[2 * i + 1] = Ops::fma(c, bb, Ops::mul(s, a));
            }
        }
    };

    if (item < Hq) {
        const int hq = item;
        T         x[4];
        load4(reinterpret_cast<const T*>(p.q) + (size_t)tok * p.stride + (size_t)hq * kHeadDim + di, x);
        if (p.q_bias) {
            T bq[4];
            load4(reinterpret_cast<const T*>(p.q_bias) + (size_t)hq * kHeadDim + di, bq);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = Ops::add(x[i], bq[i]);
        }
        rotate(x);
        T* dst = reinterpret_cast<T*>(p.q_rot) + ((size_t)tok * p.num_heads + hq) * kHeadDim + di;
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(&x[0]);
        return;
    }

    const int h = item - Hq;
    T         xk[4], xv[4];
    load4(reinterpret_cast<const T*>(p.k) + (size_t)tok * p.stride + (size_t)h * kHeadDim + di, xk);
    load4(reinterpret_cast<const T*>(p.v) + (size_t)tok * p.stride + (size_t)h * kHeadDim + di, xv);
    if (p.k_bias) {
        T bk[4], bv[4];
        load4(reinterpret_cast<const T*>(p.k_bias) + (size_t)h * kHeadDim + di, bk);
        load4(reinterpret_cast<const T*>(p.v_bias) + (size_t)h * kHeadDim + di, bv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xk[i] = Ops::add(xk[i], bk[i]);
            xv[i] = Ops::add(xv[i], bv[i]);
        }
    }
    rotate(xk);

    const int ro_len = p.readonly_block_num ? p.readonly_block_num[b] * p.block_len : 0;
    if (ti < ro_len) {
        return;
    }
    constexpr int kRow          = kHeadDim * KVBITS / 8;
    const int     head_data_sz  = p.block_len * kRow;
    const int     head_param_sz = p.block_len * 4;
    char*         blk = p.block_ptrs[p.cu_block_nums[b] + ti / p.block_len] + p.layer_offset;
    const int     bt  = ti % p.block_len;
    char*         kd  = blk + (size_t)(h * 2) * head_data_sz + (size_t)bt * kRow;
    char*         vd  = kd + head_data_sz;

    if constexpr (KVBITS == 16) {
        *reinterpret_cast<uint2*>(kd + 8 * lane) = *reinterpret_cast<uint2*>(&xk[0]);
        *reinterpret_cast<uint2*>(vd + 8 * lane) = *reinterpret_cast<uint2*>(&xv[0]);
    }
    else {
        char* kp = blk + (size_t)p.num_kv_heads * 2 * head_data_sz + (size_t)(h * 2) * head_param_sz + (size_t)bt * 4;
        char* vp = kp + head_param_sz;
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
            T(&x)[4] = kv ? xv : xk;
            float mn = Ops::to_f32(x[0]), mx = mn;
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                mn = fminf(mn, Ops::to_f32(x[i]));
                mx = fmaxf(mx, Ops::to_f32(x[i]));
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
            }
            // warp_stats (quantization.h:341-366): scale = (max - min) / (2^b - 1) in fp32, cast to T
            constexpr float inv_q_max = 1.0f / (float)((1 << KVBITS) - 1);
            const T         scale     = Ops::from_f32(__fmul_rn(__fsub_rn(mx, mn), inv_q_max));
            const T         zero      = Ops::from_f32(mn);
            const T         inv_scale = Ops::from_f32(__fdiv_rn(1.f, Ops::to_f32(scale)));
            uint32_t        q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                q[i] = Ops::round_u8(Ops::mul(Ops::sub(x[i], zero), inv_scale));
                if (KVBITS == 4) {
                    q[i] = min(q[i], 15u);
                }
            }
            char* dd = kv ? vd : kd;
            if constexpr (KVBITS == 8) {
                *reinterpret_cast<uint32_t*>(dd + 4 * lane) = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
            }
            else {
                // nibble j of each u32 holds element [0,2,4,6,1,3,5,7][j]  (ConvertKvCache<T,uint4_t>::pack)
                uint32_t mine = (lane & 1) ? ((q[0] << 8) | (q[2] << 12) | (q[1] << 24) | (q[3] << 28)) :
                                             ((q[0]) | (q[2] << 4) | (q[1] << 16) | (q[3] << 20));
                uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                if ((lane & 1) == 0) {
                    *reinterpret_cast<uint32_t*>(dd + 4 * (lane >> 1)) = mine | other;
                }
            }
            if (lane == 0) {
                T pr[2] = {scale, zero};
                *reinterpret_cast<uint32_t*>(kv ? vp : kp) = *reinterpret_cast<uint32_t*>(&pr[0]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template<int KVBITS>
constexpr int stages_for()
{
    return KVBITS == 8 ? 3 : (KVBITS == 4 ? 4 : 3);
}

template<class T, int KVBITS>
static int main_grid()
{
    constexpr int STAGES = stages_for<KVBITS>();
    using SS             = SharedStorage<T, KVBITS, STAGES>;
    auto       kern      = decode_attention_kernel<T, KVBITS, STAGES>;
    static PerDevice grids{};  // per (T, KVBITS) instantiation and per device
    const int        dev = current_device();
    int              grid = grids.value[dev].load(std::memory_order_acquire);
    if (grid == 0) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SS));
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, sizeof(SS));
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        grid = std::max(1, occ) * std::max(1, sms);
        grids.value[dev].store(grid, std::memory_order_release);
    }
    return grid;
}

static size_t align_up(size_t x, size_t a)
{
    return (x + a - 1) / a * a;
}

constexpr int kMaxGridBound = 148 * 6 * 2;  // upper bound used to size the workspace (any B200-class part)

struct Workspace {
    void*  q_rot;
    float* partial_O;
    float* partial_ML;
    int*   plan;
    int*   seg_counters;
    size_t total;
};

static Workspace carve(void* base, int token_num, int batch, int num_heads)
{
    Workspace w{};
    size_t    off = 0;
    auto      take = [&](size_t bytes) {
        void* p = base ? (char*)base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.q_rot      = take((size_t)token_num * num_heads * kHeadDim * 2);
    w.partial_O  = (float*)take((size_t)2 * kMaxGridBound * kMaxGroup * kHeadDim * 4);
    w.partial_ML = (float*)take((size_t)2 * kMaxGridBound * kMaxGroup * 2 * 4);
    w.plan       = (int*)take((size_t)(batch + 1) * 4);
    // arrival counters of the in-kernel fix-up, one per (sequence, kv-head): must be zero on first use (the kernel re-arms
    // them; the op's own plan launch clears them as well); sized for the largest kv-head count per group size 1
    w.seg_counters = (int*)take((size_t)batch * num_heads * 4);
    w.total      = off;
    return w;
}

static int validate(const b200_attention_params* a, bool decode)
{
    if (!a) {
        return fail(B200_ERR_INVALID_ARG, "params is NULL");
    }
    if (a->size_per_head != kHeadDim) {
        return fail(B200_ERR_UNSUPPORTED, "size_per_head must be 128");
    }
    if (a->window_size != 0) {
        return fail(B200_ERR_UNSUPPORTED, "window_size is not supported");
    }
    if (a->quant_policy != 0 && a->quant_policy != 4 && a->quant_policy != 8) {
        return fail(B200_ERR_INVALID_ARG, "quant_policy must be 0, 4 or 8");
    }
    if (a->dtype != B200_DTYPE_F16 && a->dtype != B200_DTYPE_BF16) {
        return fail(B200_ERR_INVALID_ARG, "dtype must be 0 (f16) or 1 (bf16)");
    }
    if (a->num_kv_heads <= 0 || a->num_heads % a->num_kv_heads) {
        return fail(B200_ERR_INVALID_ARG, "num_heads must be a multiple of num_kv_heads");
    }
    if (decode && a->num_heads / a->num_kv_heads > kMaxGroup) {
        return fail(B200_ERR_UNSUPPORTED, "GQA group size > 8 is not supported");
    }
    if (a->block_len <= 0 || a->block_len % kTile) {
        return fail(B200_ERR_UNSUPPORTED, "block_len must be a multiple of 64");
    }
    if (a->rope.type == 3) {
        return fail(B200_ERR_UNSUPPORTED, "dynamic-NTK rope is not supported");
    }
    if (a->rope.type != 0 && (a->rope.dim <= 0 || a->rope.dim > kHeadDim || a->rope.dim % 2)) {
        return fail(B200_ERR_INVALID_ARG, "rope.dim out of range");
    }
    if (decode && a->token_num != a->batch_size) {
        return fail(B200_ERR_UNSUPPORTED, "decode requires q_len == 1 for every sequence");
    }
    return 0;
}

template<class T, int KVBITS>
static int launch_prologue(const b200_attention_params* a, void* q_rot, cudaStream_t st, int wait_first = 0)
{
    PrologueParams p{};
    p.wait_first         = wait_first;
    p.q = a->q, p.k = a->k, p.v = a->v, p.stride = a->stride;
    p.q_bias = a->q_bias, p.k_bias = a->k_bias, p.v_bias = a->v_bias;
    p.q_rot              = q_rot;
    p.cu_q_len           = a->cu_q_len;
    p.cu_k_len           = a->cu_k_len;
    p.readonly_block_num = a->readonly_block_num;
    p.finished           = a->finished;
    p.block_ptrs         = (char**)a->block_ptrs;
    p.cu_block_nums      = a->cu_block_nums;
    p.layer_offset       = a->layer_offset;
    p.block_len          = a->block_len;
    p.token_num          = a->token_num;
    p.batch              = a->batch_size;
    p.num_heads          = a->num_heads;
    p.num_kv_heads       = a->num_kv_heads;
    p.rope               = make_rope_const(a->rope);
    p.rope_table         = reinterpret_cast<const float2*>(a->rope_table);
    const int items      = (q_rot ? a->num_heads : 0) + a->num_kv_heads;
    dim3      grid(a->token_num, (items + 3) / 4);
    launch_k(kv_prologue_kernel<T, KVBITS>, dim3(grid), dim3(128), 0, st, p);
    return check_launch("kv_prologue_kernel");
}

template<class T, int KVBITS>
static int launch_decode(const b200_attention_params* a)
{
    cudaStream_t st = (cudaStream_t)a->stream;
    Workspace    w  = carve(a->workspace, a->token_num, a->batch_size, a->num_heads);
    if (!a->workspace || a->workspace_bytes < w.total) {
        return fail(B200_ERR_WORKSPACE, "workspace too small (see b200_decode_attention_workspace_bytes)");
    }
    // Which parts run inside the main kernel (B200_ATTN_FUSED; debug_flags bits 2 / 3 select per call).  Default 0: the
    // three-kernel form -- prologue kernel, main kernel, fix-up kernel chained by programmatic dependent launch.  Measured
    // at C2 (profiles/r02_attention_variants.md, ncu per-kernel + CUDA events): the prologue inside the main kernel (bit 0:
    // RoPE of q in registers, the warp that owns the new token's slot quantises + stores K/V) costs the main kernel +7.2 us
    // (dependent global loads at every segment start / last tile sit on each CTA's critical path) against a 5.9 us kernel
    // that PDL overlaps; the fix-up inside (bit 1) costs +6.2 us (an acq_rel atomic round trip + two dependent L2 passes per
    // straddling segment) against an 8.5 us kernel spread over 512 CTAs.  All forms are bit-identical
    // (tests/test_decode_attention_gpu.py::test_fused_kernel_equals_the_three_kernel_form).
    static const int env_fused = [] {
        const char* e = getenv("B200_ATTN_FUSED");
        const char* u = getenv("B200_ATTN_UNFUSED");
        if (u && *u && *u != '0') {
            return 0;
        }
        return e && *e ? atoi(e) & 3 : 0;
    }();
    const int  want    = (a->debug_flags & 4) ? 0 : ((a->debug_flags & 8) ? 3 : env_fused);  // bit 3: everything in one kernel
    const bool unfused = want == 0;
    const int* plan = a->plan;
    if (!plan) {
        launch_k(plan_kernel, dim3(1), dim3(256), 0, st, w.plan, a->cu_k_len, a->finished, a->batch_size, w.seg_counters,
                 a->batch_size * a->num_kv_heads);
        if (int e = check_launch("plan_kernel")) {
            return e;
        }
        plan = w.plan;
    }
    // Early KV streaming (KernelParams::early): needs a plan computed well before this call (the caller's), the separate
    // prologue kernel, and is pointless without it.  B200_ATTN_EARLY=0 turns it off (A/B measurements).
    static const bool env_early = [] {
        const char* e = getenv("B200_ATTN_EARLY");
        return !(e && *e == '0');  // measured: 4.148 -> 4.101 ms per C2 step (profiles/r02_attention_variants.md)
    }();
    const bool run_prologue = !(want & 1) && !(a->debug_flags & 1);
    const int  early        = env_early && a->plan != nullptr && run_prologue ? 1 : 0;
    if (run_prologue) {
        if (int e = launch_prologue<T, KVBITS>(a, w.q_rot, st, early)) {
            return e;
        }
    }
    constexpr int STAGES = stages_for<KVBITS>();
    using SS             = SharedStorage<T, KVBITS, STAGES>;
    const int    grid    = main_grid<T, KVBITS>();
    if (grid > kMaxGridBound) {
        return fail(B200_ERR_WORKSPACE, "grid exceeds workspace bound");
    }
    KernelParams p{};
    p.q_rot         = w.q_rot;
    p.out           = a->out;
    p.out_stride    = (int64_t)a->num_heads * kHeadDim;
    p.block_ptrs    = (char**)a->block_ptrs;
    p.cu_block_nums = a->cu_block_nums;
    p.cu_k_len      = a->cu_k_len;
    p.tile_prefix   = plan;
    p.partial_O     = w.partial_O;
    p.partial_ML    = w.partial_ML;
    p.layer_offset  = a->layer_offset;
    p.block_len     = a->block_len;
    p.batch         = a->batch_size;
    p.num_heads     = a->num_heads;
    p.num_kv_heads  = a->num_kv_heads;
    p.qk_scale      = a->inv_sqrt_dh;
    p.out_mt        = a->out_packed_mt;
    p.grid_main     = grid;
    p.early         = early;
    if (!unfused) {
        // debug_flags bit 0 / bit 1 keep their meaning (skip the prologue / the fix-up) for kernel-only timing
        p.fused = ((a->debug_flags & 1) ? 0 : (want & 1)) | ((a->debug_flags & 2) ? 0 : (want & 2));
        p.q = a->q, p.k = a->k, p.v = a->v, p.stride = a->stride;
        p.q_bias = a->q_bias, p.k_bias = a->k_bias, p.v_bias = a->v_bias;
        p.cu_q_len           = a->cu_q_len;
        p.readonly_block_num = a->readonly_block_num;
        p.rope_table         = reinterpret_cast<const float2*>(a->rope_table);
        p.rope               = make_rope_const(a->rope);
        p.seg_counters       = w.seg_counters;
    }
    launch_k(decode_attention_kernel<T, KVBITS, STAGES>, dim3(grid), dim3(kThreads), sizeof(SS), st, p);
    if (int e = check_launch("decode_attention_kernel")) {
        return e;
    }
    if ((a->debug_flags & 2) || (want & 2)) {
        return 0;
    }
    dim3 rgrid(a->num_kv_heads, a->batch_size);
    launch_k(decode_reduce_kernel<T>, dim3(rgrid), dim3(128), 0, st, p, grid);
    return check_launch("decode_reduce_kernel");
}

// ------------------------------------------------------------------------------------------------
// prefill: prologue (bias + RoPE + KV quantise/store + rotated Q) then the tcgen05 causal attention kernel
// ------------------------------------------------------------------------------------------------
static size_t prefill_workspace_bytes(int token_num, int num_heads)
{
    return align_up((size_t)token_num * num_heads * kHeadDim * 2, 256);
}

template<class T, int KVBITS>
static int launch_prefill(const b200_attention_params* a)
{
    cudaStream_t st = (cudaStream_t)a->stream;
    if (!a->workspace || a->workspace_bytes < prefill_workspace_bytes(a->token_num, a->num_heads)) {
        return fail(B200_ERR_WORKSPACE, "workspace too small (see b200_prefill_attention_workspace_bytes)");
    }
    if (a->out_packed_mt > 0) {
        return fail(B200_ERR_UNSUPPORTED, "prefill attention writes row-major output only");
    }
    const int G = a->num_heads / a->num_kv_heads;
    if (G > kPfRows) {
        return fail(B200_ERR_UNSUPPORTED, "GQA group size > 128");
    }
    void* q_rot = a->workspace;
    if (int e = launch_prologue<T, KVBITS>(a, q_rot, st)) {
        return e;
    }
    auto kern = prefill_attention_kernel<T, KVBITS>;
    // per launch (cheap, and correct for a process that drives several devices)
    if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmem); e != cudaSuccess) {
        return fail_cuda(e, "cudaFuncSetAttribute(prefill_attention_kernel)");
    }
    PrefillParams p{};
    p.q_rot         = q_rot;
    p.out           = a->out;
    p.out_stride    = (int64_t)a->num_heads * kHeadDim;
    p.block_ptrs    = (char**)a->block_ptrs;
    p.cu_block_nums = a->cu_block_nums;
    p.cu_q_len      = a->cu_q_len;
    p.cu_k_len      = a->cu_k_len;
    p.layer_offset  = a->layer_offset;
    p.block_len     = a->block_len;
    p.batch         = a->batch_size;
    p.num_heads     = a->num_heads;
    p.num_kv_heads  = a->num_kv_heads;
    p.qk_scale      = a->inv_sqrt_dh;
    // sequence lengths live on the device: the grid is the upper bound sum_b ceil(q_len_b / TQ) <= token_num / TQ + batch,
    // surplus CTAs exit at once
    const int TQ = kPfRows / G;
    dim3      grid(a->token_num / TQ + a->batch_size, a->num_kv_heads);
    launch_k(kern, dim3(grid), dim3(kPfThreads), (size_t)kPfSmem, st, p);
    return check_launch("prefill_attention_kernel");
}

template<class F>
static int dispatch(int dtype, int quant_policy, F&& f)
{
    const int bits = quant_policy == 0 ? 16 : quant_policy;
    if (dtype == B200_DTYPE_F16) {
        if (bits == 16) return f(half{}, std::integral_constant<int, 16>{});
        if (bits == 8) return f(half{}, std::integral_constant<int, 8>{});
        return f(half{}, std::integral_constant<int, 4>{});
    }
    if (bits == 16) return f(nv_bfloat16{}, std::integral_constant<int, 16>{});
    if (bits == 8) return f(nv_bfloat16{}, std::integral_constant<int, 8>{});
    return f(nv_bfloat16{}, std::integral_constant<int, 4>{});
}

}  // namespace attn
}  // namespace b200

using namespace b200;
using namespace b200::attn;

extern "C" size_t b200_decode_attention_workspace_bytes(int batch_size, int num_heads, int num_kv_heads)
{
    (void)num_kv_heads;
    return carve(nullptr, batch_size, batch_size, num_heads).total;
}

extern "C" int b200_decode_attention_plan(int* plan, const int* cu_k_len, const uint8_t* finished, int batch_size,
                                          void* stream)
{
    if (!plan || !cu_k_len || batch_size <= 0) {
        return fail(B200_ERR_INVALID_ARG, "bad plan arguments");
    }
    launch_k(plan_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, plan, cu_k_len, finished, batch_size, (int*)nullptr, 0);
    return check_launch("plan_kernel");
}

extern "C" int b200_rope_table(float* table, const b200_rope_config* rope, const int* cu_q_len, const int* cu_k_len,
                               int batch_size, int token_num, void* stream)
{
    if (!table || !rope || !cu_q_len || !cu_k_len || batch_size <= 0 || token_num <= 0) {
        return fail(B200_ERR_INVALID_ARG, "bad rope_table arguments");
    }
    launch_k(rope_table_kernel, dim3(token_num), dim3(kHeadDim / 2), 0, (cudaStream_t)stream, (float2*)table, make_rope_const(*rope),
                                                                            cu_q_len, cu_k_len, batch_size);
    return check_launch("rope_table_kernel");
}

extern "C" int b200_decode_attention(const b200_attention_params* a)
{
    if (int e = validate(a, true)) {
        return e;
    }
    return dispatch(a->dtype, a->quant_policy, [&](auto t, auto bits) {
        using T = decltype(t);
        return launch_decode<T, decltype(bits)::value>(a);
    });
}

extern "C" int b200_process_kv(const b200_attention_params* a)
{
    if (int e = validate(a, false)) {
        return e;
    }
    return dispatch(a->dtype, a->quant_policy, [&](auto t, auto bits) {
        using T = decltype(t);
        return launch_prologue<T, decltype(bits)::value>(a, nullptr, (cudaStream_t)a->stream);
    });
}

extern "C" size_t b200_prefill_attention_workspace_bytes(int token_num, int num_heads)
{
    return prefill_workspace_bytes(token_num, num_heads);
}

extern "C" int b200_prefill_attention(const b200_attention_params* a)
{
    if (int e = validate(a, false)) {
        return e;
    }
    if (a->token_num <= 0 || a->batch_size <= 0) {
        return fail(B200_ERR_INVALID_ARG, "prefill: token_num and batch_size must be positive");
    }
    return dispatch(a->dtype, a->quant_policy, [&](auto t, auto bits) {
        using T = decltype(t);
        return launch_prefill<T, decltype(bits)::value>(a);
    });
}

extern "C" int b200_decode_attention_grid(int quant_policy, int dtype)
{
    return dispatch(dtype, quant_policy, [&](auto t, auto bits) {
        using T = decltype(t);
        return main_grid<T, decltype(bits)::value>();
    });
}
