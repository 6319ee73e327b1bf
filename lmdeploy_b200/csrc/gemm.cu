// Host side of the weight-only-quant linear layer: weight re-pack ("prepare"), activation packing and the
// launchers behind b200_linear_*.  See gemm.cuh for the kernel design.
#include "../../include/b200_turbomind.h"
#include "gemm.cuh"
#include "host_utils.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

namespace b200 {
namespace gemm {

// ------------------------------------------------------------------------------------------------
// prepare: loader-visible tensors -> tile stream consumed by gemm_kernel
//   u4 input (pre-prepare format, SURVEY 8a11 / lmdeploy/turbomind/weight_format.py:63-74):
//     qweight int32 [K, N/8], nibble i of word w = column 8w+i;  scales, zeros T [K/128, N]
//   output per (n-tile nt, k-group kg):  data [c:2][hh:2][n:128][16 B] then params [n:128]{s, t=-z*s}
//   the 32 nibbles of a 16-byte unit are k = kg*128 + c*64 + hh*32 + (0..31); inside each u32 the nibble
//   at position j holds k-offset [0,2,4,6,1,3,5,7][j] so that lop3 unpacking yields (k, k+1) pairs.
// ------------------------------------------------------------------------------------------------
template<class T>
__global__ void prepare_u4_kernel(uint8_t* dst, const uint32_t* qweight, const T* scales, const T* zeros, int K, int N)
{
    using Ops       = TypeOps<T>;
    const int kgs   = K / kKG;
    const int nts   = N / kBN;
    const size_t nw = (size_t)nts * kgs * (kU4BlockBytes / 4);  // u32 words incl. params
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nw; idx += (size_t)gridDim.x * blockDim.x) {
        const int    within = (int)(idx % (kU4BlockBytes / 4));
        const size_t blk    = idx / (kU4BlockBytes / 4);
        const int    kg     = (int)(blk % kgs);
        const int    nt     = (int)(blk / kgs);
        uint32_t     out;
        if (within < 2048) {
            const int u  = within & 3;
            const int n  = (within >> 2) & 127;
            const int hh = (within >> 9) & 1;
            const int c  = within >> 10;
            const int k0 = kg * kKG + c * 64 + hh * 32 + u * 8;
            const int ng = nt * kBN + n;
            out          = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int      ko = (j < 4) ? 2 * j : 2 * (j - 4) + 1;  // [0,2,4,6,1,3,5,7]
                const uint32_t w  = qweight[(size_t)(k0 + ko) * (N / 8) + (ng >> 3)];
                out |= ((w >> (4 * (ng & 7))) & 0xF) << (4 * j);
            }
        }
        else {
            const int n  = within - 2048;
            const int ng = nt * kBN + n;
            const T   s  = scales[(size_t)kg * N + ng];
            const T   z  = zeros ? zeros[(size_t)kg * N + ng] : Ops::from_f32(0.f);
            // fuse_scales_and_zeros (cast.cu:134-165): t = -z * s evaluated in T
            const T t = Ops::mul(Ops::from_f32(-Ops::to_f32(z)), s);
            T       pr[2] = {s, t};
            out           = *reinterpret_cast<uint32_t*>(&pr[0]);
        }
        reinterpret_cast<uint32_t*>(dst)[idx] = out;
    }
}

// u4 x f16 "subnormal" layout (gemm.cuh: kU4FastShift).  Pass 1: is every scale * 2^17 finite in f16?  Pass 2 (device-side
// decision, no host round trip): rotate the packed words left by 3, store the scales times 2^17, raise the flag.
__global__ void u4_fast_check_kernel(const uint8_t* blocks_, size_t nblocks, int* trailer)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        trailer[0] = 0;
    }
    int bad = 0;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nblocks * kBN; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t pr = reinterpret_cast<const uint32_t*>(blocks_ + (idx / kBN) * kU4BlockBytes + 8192)[idx % kBN];
        const float    sc = __half2float(__ushort_as_half((unsigned short)(pr & 0xffffu)));
        bad |= !(fabsf(sc) * (float)(1 << kU4FastShift) <= 65504.f);  // also catches NaN / inf
    }
    if (bad) {
        atomicOr(trailer + 1, 1);
    }
}

__global__ void u4_fast_apply_kernel(uint8_t* blocks_, size_t nblocks, int* trailer)
{
    if (*(volatile int*)(trailer + 1) != 0) {
        return;  // some scale is out of range: the stream keeps the plain layout, flag stays 0
    }
    const size_t nw = nblocks * (kU4BlockBytes / 4);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nw; idx += (size_t)gridDim.x * blockDim.x) {
        const int within = (int)(idx % (kU4BlockBytes / 4));
        uint32_t* w      = reinterpret_cast<uint32_t*>(blocks_) + idx;
        uint32_t  x      = *w;
        if (within < 2048) {
            x = __funnelshift_l(x, x, 3);
        }
        else {
            const float sc = __half2float(__ushort_as_half((unsigned short)(x & 0xffffu))) * (float)(1 << kU4FastShift);
            x              = (x & 0xffff0000u) | __half_as_ushort(__float2half_rn(sc));  // exact: a power-of-two multiple
        }
        *w = x;
    }
}

__global__ void u4_fast_flag_kernel(int* trailer)
{
    trailer[0] = trailer[1] == 0 ? 1 : 0;
    trailer[1] = 0;
}

// fp8 (e4m3 + 128x128 block scales, SURVEY 8a11 / linear_weight.cc:138-150): W u8 [K, N] (trans == 0) or [N, K]
// (trans == 1), scale_inv f32 [K/128, N/128]  ->  per (nt, kg): data [k16:8][n:128][16 B], then the block scale cast to T
template<class T>
__global__ void prepare_fp8_kernel(uint8_t* dst, const uint8_t* w, const float* scale_inv, int K, int N, int trans)
{
    using Ops        = TypeOps<T>;
    const int    kgs = K / kKG, nts = N / kBN;
    const size_t nb  = (size_t)nts * kgs * kFp8BlockBytes;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nb; idx += (size_t)gridDim.x * blockDim.x) {
        const int    within = (int)(idx % kFp8BlockBytes);
        const size_t blk    = idx / kFp8BlockBytes;
        const int    kg = (int)(blk % kgs), nt = (int)(blk / kgs);
        uint8_t      out = 0;
        if (within < kBN * kKG) {
            const int j = within & 15, n = (within >> 4) & 127, k16 = within >> 11;
            const int k = kg * kKG + k16 * 16 + j, ng = nt * kBN + n;
            out         = trans ? w[(size_t)ng * K + k] : w[(size_t)k * N + ng];
        }
        else if (within < kBN * kKG + 2) {
            const T s = Ops::from_f32(scale_inv[(size_t)kg * nts + nt]);
            out       = reinterpret_cast<const uint8_t*>(&s)[within - kBN * kKG];
        }
        dst[idx] = out;
    }
}

// fp8 with per-column scales: same data layout; scales f32 [K/128, N] -> one T per row of the block (4-byte stride)
template<class T>
__global__ void prepare_fp8col_kernel(uint8_t* dst, const uint8_t* w, const float* scales, int K, int N, int trans)
{
    using Ops        = TypeOps<T>;
    const int    kgs = K / kKG, nts = N / kBN;
    const size_t nb  = (size_t)nts * kgs * kFp8ColBlockBytes;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nb; idx += (size_t)gridDim.x * blockDim.x) {
        const int    within = (int)(idx % kFp8ColBlockBytes);
        const size_t blk    = idx / kFp8ColBlockBytes;
        const int    kg = (int)(blk % kgs), nt = (int)(blk / kgs);
        uint8_t      out = 0;
        if (within < kBN * kKG) {
            const int j = within & 15, n = (within >> 4) & 127, k16 = within >> 11;
            const int k = kg * kKG + k16 * 16 + j, ng = nt * kBN + n;
            out         = trans ? w[(size_t)ng * K + k] : w[(size_t)k * N + ng];
        }
        else {
            const int rel = within - kBN * kKG, n = rel >> 2, b = rel & 3;
            if (b < 2) {
                const T s = Ops::from_f32(scales[(size_t)kg * N + nt * kBN + n]);
                out       = reinterpret_cast<const uint8_t*>(&s)[b];
            }
        }
        dst[idx] = out;
    }
}

// dense: W [K, N] row-major (trans == 0) or [N, K] row-major (trans == 1) -> per (nt, kg): [k8:16][n:128][8] T
template<class T>
__global__ void prepare_dense_kernel(T* dst, const T* w, int K, int N, int trans)
{
    const int    kgs = K / kKG;
    const size_t ne  = (size_t)K * N;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < ne; idx += (size_t)gridDim.x * blockDim.x) {
        const int    e   = (int)(idx & 7);
        const int    n   = (int)((idx >> 3) & 127);
        const int    k8  = (int)((idx >> 10) & 15);
        const size_t blk = idx >> 14;
        const int    kg  = (int)(blk % kgs);
        const int    nt  = (int)(blk / kgs);
        const int    k   = kg * kKG + k8 * 8 + e;
        const int    ng  = nt * kBN + n;
        dst[idx]         = trans ? w[(size_t)ng * K + k] : w[(size_t)k * N + ng];
    }
}

// activations: X [M][ldx] row-major -> XP [K/8][MT][8]; rows >= M are zero
template<class T>
__global__ void pack_x_kernel(T* xp, const T* x, int M, int MT, int K, int64_t ldx)
{
    const size_t n16 = (size_t)(K / 8) * MT;  // 16-byte units
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n16; idx += (size_t)gridDim.x * blockDim.x) {
        const int tok = (int)(idx % MT);
        const int k8  = (int)(idx / MT);
        uint4     v   = make_uint4(0, 0, 0, 0);
        if (tok < M) {
            v = *reinterpret_cast<const uint4*>(x + (size_t)tok * ldx + (size_t)k8 * 8);
        }
        reinterpret_cast<uint4*>(xp)[idx] = v;
    }
}

constexpr int kMaxSplits = 16;

struct Plan {
    int tiles, splits, nw, nx, na, tmem_cols, a_col, smem;
};

// Launch plan: a pure function of (format, shape, padded tokens) => deterministic results.
//   * u4 with N % 256 == 0: two n-tiles per CTA (shared activation tile), ONE CTA per SM with the whole
//     shared-memory / TMEM budget spent on ring depth (the pipeline, not a second CTA, hides the latency);
//   * otherwise one n-tile per CTA (N not a multiple of 256, dense weights), also one CTA per SM;
//   * K is split over 1/2/4/8 co-resident CTAs (largest count that fits one wave, every CTA >= 4 K-groups).
// A plan the heuristic below is told to use instead of its own choice: fields > 0 replace the heuristic's value.
struct PlanOverride {
    int tiles, splits, nw, nx;
    int variant;  // tool builds only (gemm.cuh Params::variant)
};

#ifdef B200_GEMM_TOOL
// tools/gemm_sweep.cu compiles this TU with -DB200_GEMM_TOOL to measure the plan space.  Never compiled into the
// product library.
static PlanOverride g_plan_override{};
#endif

// Tuned plans per (format, padded tokens, N, K): the counterpart of the reference's GEMM dispatch cache
// (kernels/gemm/dispatch_cache.cu; Gemm::Export / Gemm::Import, gemm.h:43-45).  Filled by b200_linear_load_plans from
// a text file that tools/gemm_sweep.cu can write ("PLAN fmt mt N K tiles splits nw nx" lines); empty = heuristic only.
struct PlanKey {
    int fmt, mt, N, K;
    bool operator<(const PlanKey& o) const
    {
        return std::tie(fmt, mt, N, K) < std::tie(o.fmt, o.mt, o.N, o.K);
    }
};
static std::map<PlanKey, PlanOverride> g_plan_table;
static std::mutex                      g_plan_mutex;

static bool lookup_plan(int fmt, int MT, int N, int K, PlanOverride* out)
{
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto                        it = g_plan_table.find(PlanKey{fmt, MT, N, K});
    if (it == g_plan_table.end()) {
        return false;
    }
    *out = it->second;
    return true;
}

static Plan make_plan(int fmt, int MT, int N, int K, bool need_split = false)
{
    const int WB     = block_bytes_of(fmt);
    const int WB_pad = (WB + 1023) & ~1023;
    const int XB     = MT * 256;
    Plan      pl{};
    pl.tiles         = (fmt != kDense && N % (2 * kBN) == 0) ? 2 : 1;
    PlanOverride ov{};
    lookup_plan(fmt, MT, N, K, &ov);
#ifdef B200_GEMM_TOOL
    if (g_plan_override.tiles > 0 || g_plan_override.splits > 0 || g_plan_override.nw > 0 || g_plan_override.variant) {
        ov = g_plan_override;
    }
#endif
    if (ov.tiles > 0 && fmt != kDense && N % (kBN * ov.tiles) == 0 && (ov.tiles <= 2 || MT <= 64)) {
        pl.tiles = ov.tiles;  // 4 tiles (tool builds only): 4 x 64 accumulator columns, so <= 64 tokens
    }
    const bool two_per_sm = false;  // 640-thread CTAs at ~90 registers: one CTA per SM in every configuration
    const int  nts = N / (kBN * pl.tiles), kgs = K / kKG;
    const int  slots = 148 * (two_per_sm ? 2 : 1);
    // K splits: largest power of two (<= 8) that fits one wave and leaves every CTA >= 4 K-groups.  (Measured: 6 / 9
    // splits fill 144 of 148 SMs but are slower -- the exchange traffic grows faster than the tail shrinks.)
    pl.splits = 1;
    while (pl.splits < 8 && nts * (pl.splits * 2) <= slots && kgs / (pl.splits * 2) >= 4) {
        pl.splits *= 2;
    }
    // Short K (<= 16 groups: the row-split wo / w2 of a tensor-parallel shard): one tile per CTA and NO split beats any
    // split plan -- the 4.3 us exchange costs more than the K-groups it saves (measured on the per-rank shapes of
    // InternLM2-7B / 20B and Llama-3-70B at TP 2-8, profiles/r02_raw/gemm_sweep_*: e.g. K 1024, N 4096: 11.8 -> 5.3 us).
    // (need_split: the fused residual + RMSNorm epilogue lives in the split-K exchange and keeps the split plan.)
    if (!need_split && fmt != kDense && ov.tiles == 0 && ov.splits == 0 && kgs <= 16 && N / kBN <= slots) {
        pl.tiles  = 1;
        pl.splits = 1;
    }
    if (ov.splits > 0) {
        pl.splits = ov.splits;
    }
    const int budget = (two_per_sm ? 113 : 224) * 1024 - 1024;
    const int wslot  = WB_pad * pl.tiles;
    // both rings want ~5-6 slots in flight (each loop is ~2.5-3k cycles of latency per slot)
    pl.nx = std::min(kMaxX, std::max(3, (budget * 2 / 5) / XB));
    pl.nw = std::min(kMaxW, (budget - pl.nx * XB) / wslot);
    if (pl.nw < 3) {
        pl.nw = 3;
        pl.nx = std::max(2, std::min(kMaxX, (budget - pl.nw * wslot) / XB));
    }
    // u4: a W ring of exactly 8 slots selects the de-quantisation loop that is unrolled over one lap of the ring (all ring
    // positions compile-time constants, 126 instead of 162 instructions per k-group); trade X slots for it while >= 3 remain
    if (fmt == kU4 && pl.nw < 8) {
        for (int nx = pl.nx - 1; nx >= 3; --nx) {
            if ((budget - nx * XB) / wslot >= 8) {
                pl.nx = nx, pl.nw = 8;
                break;
            }
        }
    }
    if (ov.nw > 0 && ov.nx > 0) {
        pl.nw = std::min(kMaxW, ov.nw), pl.nx = std::min(kMaxX, ov.nx);
    }
    pl.smem = pl.nw * wslot + pl.nx * XB;
    const int d_cols = (MT <= 64 ? 64 : 128) * pl.tiles;
    if (fmt == kDense) {
        pl.tmem_cols = MT <= 32 ? 32 : (MT <= 64 ? 64 : 128);
        pl.a_col = 0, pl.na = 1;
    }
    else if (two_per_sm) {
        pl.tmem_cols = 256, pl.a_col = 64, pl.na = 3;
    }
    else {
        pl.tmem_cols = 512, pl.a_col = d_cols;
        pl.na        = std::min(kASlots, (512 - d_cols) / 64);
        pl.na -= pl.na % pl.tiles;
    }
    return pl;
}

constexpr size_t kCounterBytes = 64 * 1024;  // 2 ints per n-group, N <= 1M

struct LinearWs {
    int*   counters;
    int*   norm_ctr;
    void*  xp;
    float* partials;
    float* ssq;
    size_t total;
};

static size_t align_up(size_t x, size_t a)
{
    return (x + a - 1) / a * a;
}

static LinearWs carve_ws(void* base, int K, int N, int MT)
{
    LinearWs w{};
    size_t   off  = 0;
    auto     take = [&](size_t bytes) {
        void* p = base ? (char*)base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    // fixed-size counter block so that linears of different shapes can share one workspace: a counter slot never
    // aliases another linear's activation / partial region
    w.counters = (int*)take(kCounterBytes);
    w.norm_ctr = w.counters ? w.counters + kCounterBytes / 8 : nullptr;  // second half of the counter block
    w.xp       = take((size_t)K * MT * 2);
    w.partials = (float*)take((size_t)N * kMaxSplits * MT * 4);  // up to kMaxSplits K-splits of an [MT][N] fp32 tile
    w.ssq      = (float*)take((size_t)(N / kBN) * 128 * 4);      // fused RMSNorm: per n-group partial sums of squares
    w.total    = off;
    return w;
}

template<class T, int FMT, int TILES>
static int launch(const Params& p, int smem, cudaStream_t st)
{
    auto        kern       = gemm_kernel<T, FMT, TILES>;
    static PerDevice configured{};  // per instantiation and per device
    const int        dev = current_device();
    if (configured.value[dev].load(std::memory_order_acquire) == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        if (e != cudaSuccess) {
            return fail_cuda(e, "cudaFuncSetAttribute(gemm_kernel)");
        }
        configured.value[dev].store(1, std::memory_order_release);
    }
    if (smem > 227 * 1024 - 2048) {
        return fail(B200_ERR_UNSUPPORTED, "gemm: shared memory budget exceeded");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = dim3(p.N / (kBN * TILES), p.splits, 1);
    cfg.blockDim         = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = st;
    // split-K CTAs wait for each other (epilogue part 2): the launch must guarantee co-residency.
    // PDL (host_utils.h): the kernel prefetches weights before griddepcontrol.wait.
    cudaLaunchAttribute attr[2];
    int                 na = 0;
    if (p.splits > 1) {
        attr[na].id              = cudaLaunchAttributeCooperative;
        attr[na].val.cooperative = 1;
        ++na;
    }
    attr[na].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    ++na;
    cfg.attrs    = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) {
        return fail_cuda(e, "cudaLaunchKernelEx(gemm_kernel)");
    }
    return check_launch("gemm_kernel");
}

}  // namespace gemm
}  // namespace b200

using namespace b200;
using namespace b200::gemm;

static int validate_desc(const b200_linear_desc* d)
{
    if (!d) {
        return fail(B200_ERR_INVALID_ARG, "linear desc is NULL");
    }
    if (d->input_dim <= 0 || d->input_dim % kKG) {
        return fail(B200_ERR_UNSUPPORTED, "input_dim must be a multiple of 128");
    }
    if (d->output_dim <= 0 || d->output_dim % kBN || (size_t)(d->output_dim / kBN) * 8 > kCounterBytes / 2) {
        return fail(B200_ERR_UNSUPPORTED, "output_dim must be a multiple of 128 (and <= 1M)");
    }
    if (d->weight_format == kU4 && d->group_size != 128) {
        return fail(B200_ERR_UNSUPPORTED, "u4 weights require group_size 128");
    }
    if (d->weight_format != kU4 && d->weight_format != kDense && d->weight_format != kFp8 && d->weight_format != kFp8Col) {
        return fail(B200_ERR_UNSUPPORTED, "weight_format must be 0 (dense), 1 (u4 awq), 2 (fp8 e4m3, 128x128 block scales) or "
                                          "3 (fp8 e4m3, per-column scales of 128-k groups)");
    }
    if ((d->weight_format == kFp8 || d->weight_format == kFp8Col) && d->group_size != 128) {
        return fail(B200_ERR_UNSUPPORTED, "fp8 weights require 128-wide scale groups along K");
    }
    if (d->dtype != B200_DTYPE_F16 && d->dtype != B200_DTYPE_BF16) {
        return fail(B200_ERR_INVALID_ARG, "dtype must be 0 or 1");
    }
    return 0;
}

extern "C" size_t b200_linear_prepared_bytes(const b200_linear_desc* d)
{
    if (validate_desc(d)) {
        return 0;
    }
    const size_t blocks = (size_t)(d->output_dim / kBN) * (d->input_dim / kKG);
    return blocks * block_bytes_of(d->weight_format) + (d->weight_format == kU4 ? kU4TrailerBytes : 0);
}

static size_t u4_trailer_offset(const b200_linear_desc* d)
{
    return (size_t)(d->output_dim / kBN) * (d->input_dim / kKG) * kU4BlockBytes;
}

extern "C" int b200_linear_prepare(const b200_linear_desc* d, const void* weight, const void* scales, const void* zeros,
                                   int weight_is_transposed, void* prepared, void* stream)
{
    if (int e = validate_desc(d)) {
        return e;
    }
    if (!weight || !prepared) {
        return fail(B200_ERR_INVALID_ARG, "weight / prepared is NULL");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int    K = d->input_dim, N = d->output_dim;
    if (d->weight_format == kU4) {
        if (!scales) {
            return fail(B200_ERR_INVALID_ARG, "u4 weights need scales");
        }
        int* trailer = reinterpret_cast<int*>((uint8_t*)prepared + u4_trailer_offset(d));
        cudaMemsetAsync(trailer, 0, kU4TrailerBytes, st);
        if (d->dtype == B200_DTYPE_F16) {
            prepare_u4_kernel<half><<<1024, 256, 0, st>>>((uint8_t*)prepared, (const uint32_t*)weight,
                                                          (const half*)scales, (const half*)zeros, K, N);
            static const bool plain = [] {
                const char* e = getenv("B200_U4_PLAIN");  // keep the plain layout (A/B measurements, tests of the fallback)
                return e && *e && *e != '0';
            }();
            if (!plain) {
                const size_t nblocks = (size_t)(N / kBN) * (K / kKG);
                u4_fast_check_kernel<<<512, 256, 0, st>>>((const uint8_t*)prepared, nblocks, trailer);
                u4_fast_apply_kernel<<<1024, 256, 0, st>>>((uint8_t*)prepared, nblocks, trailer);
                u4_fast_flag_kernel<<<1, 1, 0, st>>>(trailer);
            }
        }
        else {
            prepare_u4_kernel<nv_bfloat16><<<1024, 256, 0, st>>>((uint8_t*)prepared, (const uint32_t*)weight,
                                                                 (const nv_bfloat16*)scales,
                                                                 (const nv_bfloat16*)zeros, K, N);
        }
    }
    else if (d->weight_format == kFp8) {
        if (!scales) {
            return fail(B200_ERR_INVALID_ARG, "fp8 weights need the f32 block scales (weight_scale_inv)");
        }
        if (d->dtype == B200_DTYPE_F16) {
            prepare_fp8_kernel<half><<<2048, 256, 0, st>>>((uint8_t*)prepared, (const uint8_t*)weight, (const float*)scales,
                                                           K, N, weight_is_transposed);
        }
        else {
            prepare_fp8_kernel<nv_bfloat16><<<2048, 256, 0, st>>>((uint8_t*)prepared, (const uint8_t*)weight,
                                                                  (const float*)scales, K, N, weight_is_transposed);
        }
    }
    else if (d->weight_format == kFp8Col) {
        if (!scales) {
            return fail(B200_ERR_INVALID_ARG, "fp8 weights need the f32 scales [K/128, N]");
        }
        if (d->dtype == B200_DTYPE_F16) {
            prepare_fp8col_kernel<half><<<2048, 256, 0, st>>>((uint8_t*)prepared, (const uint8_t*)weight,
                                                              (const float*)scales, K, N, weight_is_transposed);
        }
        else {
            prepare_fp8col_kernel<nv_bfloat16><<<2048, 256, 0, st>>>((uint8_t*)prepared, (const uint8_t*)weight,
                                                                     (const float*)scales, K, N, weight_is_transposed);
        }
    }
    else {
        prepare_dense_kernel<uint16_t><<<2048, 256, 0, st>>>((uint16_t*)prepared, (const uint16_t*)weight, K, N,
                                                             weight_is_transposed);
    }
    return check_launch("prepare kernel");
}

extern "C" size_t b200_linear_workspace_bytes(const b200_linear_desc* d, int max_tokens)
{
    if (validate_desc(d)) {
        return 0;
    }
    const int MT = std::max(16, (max_tokens + 15) / 16 * 16);
    return carve_ws(nullptr, d->input_dim, d->output_dim, MT).total;
}

extern "C" int b200_pack_activations(const void* x, int64_t ldx, int m, int k, int mt, int dtype, void* xp, void* stream)
{
    (void)dtype;
    if (!x || !xp || m < 0 || mt < m || mt % 16 || k % 8) {
        return fail(B200_ERR_INVALID_ARG, "bad pack arguments");
    }
    const size_t n16 = (size_t)(k / 8) * mt;
    const int    blocks = (int)std::min<size_t>((n16 + 255) / 256, 4096);
    pack_x_kernel<uint16_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((uint16_t*)xp, (const uint16_t*)x, m, mt, k, ldx);
    return check_launch("pack_x_kernel");
}

struct NormFusion {
    void*       residual;
    const void* weight;
    float       eps;
    float*      ssq;
    int*        ctr;
};

static int forward_packed(const b200_linear_desc* d, const void* xp, int mt, int m, void* y, int64_t ldy, int epilogue,
                          int out_mode, int out_mt, float* partials, int* counters, cudaStream_t st,
                          const NormFusion* nf = nullptr)
{
    if (mt < 16 || mt > 128 || mt % 16) {
        return fail(B200_ERR_UNSUPPORTED, "padded token count must be a multiple of 16 in [16, 128]");
    }
    if (epilogue != 0 && epilogue != 1 && !(epilogue == 2 && nf)) {
        return fail(B200_ERR_INVALID_ARG, "epilogue must be 0 (none) or 1 (gated SiLU)");
    }
    Params p{};
    p.w        = (const uint8_t*)d->prepared;
    p.u4_fast  = d->weight_format == kU4 ? reinterpret_cast<const int*>((const uint8_t*)d->prepared + u4_trailer_offset(d)) : nullptr;
    p.xp       = xp;
    p.y        = y;
    p.bias     = d->bias;
    p.partials = partials;
    p.counters = counters;
    p.N        = d->output_dim;
    p.K        = d->input_dim;
    p.MT       = mt;
    p.M        = m;
    const Plan rg = make_plan(d->weight_format, mt, p.N, p.K, epilogue == 2);
    p.splits = rg.splits, p.nw = rg.nw, p.nx = rg.nx, p.na = rg.na;
#ifdef B200_GEMM_TOOL
    p.variant = g_plan_override.variant;
#endif
#ifdef B200_GEMM_TRACE
    static int trace_launch = 0;
    p.trace_launch          = trace_launch++;
#endif
    p.tmem_cols = rg.tmem_cols, p.a_col = rg.a_col;
    if (epilogue == 2) {
        // the fused residual + RMSNorm epilogue lives in the split-K exchange: it needs splits > 1 and a bounded number
        // of (token, column) items per thread; callers fall back to the separate kernel otherwise
        const int ntok_max = (mt + rg.splits - 1) / rg.splits;
        if (rg.splits == 1 || ntok_max * (kBN * rg.tiles / 2) > kNormMaxItems * kDqWarps * 32 || ntok_max > 32) {
            return fail(B200_ERR_UNSUPPORTED, "fused residual+RMSNorm epilogue not available for this shape");
        }
        p.residual = nf->residual, p.norm_w = nf->weight, p.norm_eps = nf->eps, p.ssq = nf->ssq, p.norm_ctr = nf->ctr;
    }
    p.epilogue = epilogue;
    p.out_mode = out_mode;
    p.ldy      = (int)ldy;
    p.out_mt   = out_mt;
    const bool bf = d->dtype == B200_DTYPE_BF16;
    if (d->weight_format == kU4) {
#ifdef B200_GEMM_TOOL
        if (rg.tiles == 4) {
            return bf ? launch<nv_bfloat16, kU4, 4>(p, rg.smem, st) : launch<half, kU4, 4>(p, rg.smem, st);
        }
#endif
        if (rg.tiles == 2) {
            return bf ? launch<nv_bfloat16, kU4, 2>(p, rg.smem, st) : launch<half, kU4, 2>(p, rg.smem, st);
        }
        return bf ? launch<nv_bfloat16, kU4, 1>(p, rg.smem, st) : launch<half, kU4, 1>(p, rg.smem, st);
    }
    if (d->weight_format == kFp8) {
        if (rg.tiles == 2) {
            return bf ? launch<nv_bfloat16, kFp8, 2>(p, rg.smem, st) : launch<half, kFp8, 2>(p, rg.smem, st);
        }
        return bf ? launch<nv_bfloat16, kFp8, 1>(p, rg.smem, st) : launch<half, kFp8, 1>(p, rg.smem, st);
    }
    if (d->weight_format == kFp8Col) {
        if (rg.tiles == 2) {
            return bf ? launch<nv_bfloat16, kFp8Col, 2>(p, rg.smem, st) : launch<half, kFp8Col, 2>(p, rg.smem, st);
        }
        return bf ? launch<nv_bfloat16, kFp8Col, 1>(p, rg.smem, st) : launch<half, kFp8Col, 1>(p, rg.smem, st);
    }
    return bf ? launch<nv_bfloat16, kDense, 1>(p, rg.smem, st) : launch<half, kDense, 1>(p, rg.smem, st);
}

extern "C" int b200_linear_forward(const b200_linear_desc* d, const void* x, int64_t ldx, int m, void* y, int64_t ldy,
                                   int epilogue, void* workspace, size_t workspace_bytes, void* stream)
{
    if (int e = validate_desc(d)) {
        return e;
    }
    if (!d->prepared || !x || !y) {
        return fail(B200_ERR_INVALID_ARG, "prepared / x / y is NULL");
    }
    if (m <= 0 || m > 128) {
        return fail(B200_ERR_UNSUPPORTED, "m must be in [1, 128] (decode batch)");
    }
    const int mt = std::max(16, (m + 15) / 16 * 16);
    LinearWs  w  = carve_ws(workspace, d->input_dim, d->output_dim, mt);
    if (!workspace || workspace_bytes < w.total) {
        return fail(B200_ERR_WORKSPACE, "workspace too small (see b200_linear_workspace_bytes)");
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = b200_pack_activations(x, ldx, m, d->input_dim, mt, d->dtype, w.xp, stream)) {
        return e;
    }
    return forward_packed(d, w.xp, mt, m, y, ldy, epilogue, 0, 0, w.partials, w.counters, st);
}

extern "C" int b200_linear_forward_packed(const b200_linear_desc* d, const void* xp, int mt, int m, void* y,
                                          int64_t ldy, int epilogue, int out_packed_mt, void* workspace,
                                          size_t workspace_bytes, void* stream)
{
    if (int e = validate_desc(d)) {
        return e;
    }
    if (!d->prepared || !xp || !y) {
        return fail(B200_ERR_INVALID_ARG, "prepared / xp / y is NULL");
    }
    LinearWs w = carve_ws(workspace, d->input_dim, d->output_dim, mt);
    if (!workspace || workspace_bytes < w.total) {
        return fail(B200_ERR_WORKSPACE, "workspace too small (see b200_linear_workspace_bytes)");
    }
    return forward_packed(d, xp, mt, m, y, ldy, epilogue, out_packed_mt > 0 ? 1 : 0, out_packed_mt, w.partials,
                          w.counters, (cudaStream_t)stream);
}

extern "C" int b200_linear_forward_packed_residual_rmsnorm(const b200_linear_desc* d, const void* xp, int mt, int m,
                                                           void* residual, const void* norm_weight, float eps, void* out,
                                                           int out_packed_mt, void* workspace, size_t workspace_bytes,
                                                           void* stream)
{
    if (int e = validate_desc(d)) {
        return e;
    }
    if (!d->prepared || !xp || !residual || !norm_weight || !out) {
        return fail(B200_ERR_INVALID_ARG, "prepared / xp / residual / norm_weight / out is NULL");
    }
    LinearWs w = carve_ws(workspace, d->input_dim, d->output_dim, mt);
    if (!workspace || workspace_bytes < w.total) {
        return fail(B200_ERR_WORKSPACE, "workspace too small (see b200_linear_workspace_bytes)");
    }
    NormFusion nf{residual, norm_weight, eps, w.ssq, w.norm_ctr};
    return forward_packed(d, xp, mt, m, out, d->output_dim, 2, out_packed_mt > 0 ? 1 : 0, out_packed_mt, w.partials,
                          w.counters, (cudaStream_t)stream, &nf);
}

// Host-only introspection of the launch plan (tests/test_linear_plan_cpu.py checks its invariants without a GPU).
// Checks one tuned plan against the budgets the kernel relies on (the same ones tests/test_linear_plan_cpu.py asserts).
static bool plan_entry_ok(const PlanKey& k, const PlanOverride& v)
{
    if (k.fmt < kDense || k.fmt > kFp8Col || k.mt < 16 || k.mt > 128 || k.mt % 16 || k.N <= 0 || k.N % kBN || k.K <= 0
        || k.K % kKG) {
        return false;
    }
    if (v.tiles < 1 || v.tiles > 2 || (k.fmt == kDense && v.tiles != 1) || k.N % (kBN * v.tiles)) {
        return false;
    }
    const int nts = k.N / (kBN * v.tiles), kgs = k.K / kKG;
    if (v.splits < 1 || v.splits > kMaxSplits || v.splits > kgs || (v.splits > 1 && nts * v.splits > 148)) {
        return false;
    }
    if ((v.nw > 0) != (v.nx > 0)) {
        return false;
    }
    if (v.nw > 0) {
        const int WB_pad = (block_bytes_of(k.fmt) + 1023) & ~1023;
        if (v.nw < 3 || v.nw > kMaxW || v.nx < 2 || v.nx > kMaxX || v.nw * WB_pad * v.tiles + v.nx * k.mt * 256 > 223 * 1024) {
            return false;
        }
    }
    return true;
}

extern "C" int b200_linear_load_plans(const char* path)
{
    if (!path) {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        g_plan_table.clear();
        return 0;
    }
    FILE* f = fopen(path, "r");
    if (!f) {
        return fail(B200_ERR_INVALID_ARG, "linear_load_plans: cannot open the plan file");
    }
    std::map<PlanKey, PlanOverride> table;
    char                            line[256];
    int                             lineno = 0;
    while (fgets(line, sizeof(line), f)) {
        ++lineno;
        PlanKey      k{};
        PlanOverride v{};
        if (line[0] == '#' || line[0] == '\n') {
            continue;
        }
        if (sscanf(line, "PLAN %d %d %d %d %d %d %d %d", &k.fmt, &k.mt, &k.N, &k.K, &v.tiles, &v.splits, &v.nw, &v.nx) != 8
            || !plan_entry_ok(k, v)) {
            fclose(f);
            last_error_ref() = "linear_load_plans: invalid entry at line " + std::to_string(lineno);
            return B200_ERR_INVALID_ARG;
        }
        table[k] = v;
    }
    fclose(f);
    const int n = (int)table.size();
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        g_plan_table.swap(table);
    }
    return n;
}

extern "C" int b200_linear_plan(const b200_linear_desc* d, int m, int* out8)
{
    if (int e = validate_desc(d)) {
        return e;
    }
    if (!out8 || m <= 0 || m > 128) {
        return fail(B200_ERR_INVALID_ARG, "linear_plan: bad arguments");
    }
    const int  mt = std::max(16, (m + 15) / 16 * 16);
    const Plan pl = make_plan(d->weight_format, mt, d->output_dim, d->input_dim);
    out8[0] = pl.tiles, out8[1] = pl.splits, out8[2] = pl.nw, out8[3] = pl.nx, out8[4] = pl.na;
    out8[5] = pl.tmem_cols, out8[6] = pl.a_col, out8[7] = pl.smem;
    return 0;
}
