// Weight-only-quant GEMM for decode (M = tokens <= 128) on tcgen05 tensor cores, sm_100a.
//
// Replaces, behind the C ABI, LlamaLinear::Forward -> gemm::Gemm::Run
//   (src/turbomind/models/llama/LlamaLinear.cu:140-216, kernels/gemm/gemm.cu:257-344) together with the
//   paired weight re-packer LinearWeight::prepare (models/linear_weight.cc:101-324).
// It is NOT a port of the sm80 mma.sync kernels: the problem is computed transposed,
//     Y^T[N_out, tokens] = W^T[N_out, K] . X^T[K, tokens]
// so the streamed operand (weights) sits on the UMMA M side (128 output features per CTA) and the tiny
// activation tile is UMMA N (16..128 tokens).  Per 128-wide K group (= one AWQ quant group):
//   * warp 0 streams the packed weights (one contiguous 8.5 KB 1-D TMA bulk copy per (n-tile, group): u4 data +
//     fused (scale, -zero*scale) pairs; fp8: 16 KB e4m3 + block scale) into a deep W ring; warp 18 streams the
//     activation tiles (already in the UMMA K-major "interleaved" layout in global memory, L2 resident) into a
//     separate X ring.  The rings are decoupled because their lifetimes differ: a W slot is free as soon as it has
//     been de-quantised (long before its MMAs retire) and it is the DRAM-latency-bound stream, so it gets the depth.
//     A CTA normally owns TWO n-tiles (256 output features) that share every activation tile;
//   * 16 dequant warps (one group of 8 per n-tile) unpack u4 -> f16 with lop3 and apply w = fma(q, s, t) in f16
//     (bit-identical to the reference's transform, transform.h:48-74 / cast.cu:155; fp8: T(e4m3) *_T T(scale)) and
//     write the A operand straight into TENSOR MEMORY with tcgen05.st (ring of 64-column slots; no shared-memory
//     round trip for the dequantised weights); the loop is software-pipelined (next group's LDS under this group's st);
//   * warps 1 and 19 (one per n-tile) issue tcgen05.mma (A from TMEM, B from smem) from a warp-uniform loop -- one
//     elected lane, descriptors in uniform registers -- fp32 accumulators in TMEM;
//   * K is split over grid.y (co-resident CTAs, cooperative launch); the partial accumulators are reduce-scattered
//     by token through an L2-resident fp32 exchange buffer and summed in rank order (deterministic), then bias /
//     gated-SiLU and the fp16/bf16 store, all inside the same kernel.  (A thread-block-cluster + DSMEM exchange
//     was measured first: correct, but launching 8-CTA clusters of full-SM CTAs costs ~20 us per kernel.)
// The kernel is HBM-bound at decode batch sizes: algorithmic bytes = packed weight stream, read once.
#pragma once

#include "umma.cuh"

#include <type_traits>

namespace b200 {
namespace gemm {

constexpr int kBN        = 128;  // output features per CTA (UMMA M)
constexpr int kKG        = 128;  // K per pipeline stage (= quant group size)
constexpr int kNormMaxItems = 4;  // fused residual+RMSNorm epilogue: (token, column-pair) items per thread
constexpr int kMaxW     = 8;    // W ring slots (barrier arrays)
constexpr int kMaxX     = 6;    // X ring slots
constexpr int kASlots   = 4;    // max TMEM A-operand ring slots (64 columns each); Params::na in use
constexpr int kDqWarps  = 16;   // dequant warps: 4 per TMEM lane quarter, one 32-wide k-chunk each
constexpr int kXWarp    = 2 + kDqWarps;
constexpr int kMmaWarp2 = kXWarp + 1;  // second MMA issuer (tile 1 of a two-tile CTA)
constexpr int kThreads  = 32 * (kMmaWarp2 + 1);  // warp0 W-TMA, warp1 MMA/TMEM, warps2-17 dequant (+ epilogue), warp18 X-TMA,
                                                 // warp19 MMA of tile 1
// TMEM budget (one CTA per SM): D accumulators of tile t at column t * d_stride (64 or 128 columns each), then the
// A-operand ring of Params::na 64-column slots starting at column a_col (gemm.cu: make_plan).

enum WeightFormat : int
{
    kDense = 0,  // T weights, [k/8][128][8] tiles
    kU4    = 1,  // AWQ-style u4 + (scale, zero) per 128-group
    kFp8   = 2,  // e4m3 + per-128x128-block scale
    kFp8Col = 3, // e4m3 + per-(128-k-group, output column) scale: what the reference's kernel consumes after
                 // LinearWeight::prepare expands the checkpoint's block scales (linear_weight.cc:138-150); lets the
                 // loader permute / interleave / slice output columns freely (RoPE reorder, fused gate|up, TP)
};

constexpr int kU4BlockBytes    = 8192 + 512;           // one (n-tile, k-group): data + fused params
// u4 x f16 "subnormal" layout (chosen by b200_linear_prepare when every |scale| * 2^17 is finite in f16, i.e. always for real
// checkpoints): a nibble q sitting at bits 7..10 of an f16 lane IS the number q * 2^-17 (the f16 encoding is linear across the
// subnormals and the first binade), so  fma(q * 2^-17, s * 2^17, t)  ==  fma(q, s, t)  bit for bit -- the same single rounding
// of the same real number -- without the magic-number subtraction: 11 instead of 13 instructions per 8 weights, and 4 instead of
// 8 of them on the fma pipe.  The packed words are stored rotated left by 3 so that the nibble fields sit at bits 3, 7, .. 31
// (the last one wraps) and three shifts / one rotate bring every pair to bits 7..10 / 23..26; scales are stored times 2^17.
constexpr int      kU4FastShift  = 17;
constexpr uint32_t kU4FastMask   = 0x07800780u;
constexpr int      kU4TrailerBytes = 256;  // after the last block: word 0 = layout flag, word 1 = scratch of the range check
constexpr int kDenseBlockBytes = kBN * kKG * 2;        // 32 KB
constexpr int kFp8BlockBytes   = kBN * kKG + 16;       // e4m3 data [k16:8][n:128][16 B] + block scale (T, padded to 16 B)
constexpr int kFp8ColBlockBytes = kBN * kKG + kBN * 4; // same data + one scale per row (T in the low half of 4 B)

constexpr int block_bytes_of(int fmt)
{
    return fmt == kU4 ? kU4BlockBytes : fmt == kDense ? kDenseBlockBytes : fmt == kFp8 ? kFp8BlockBytes : kFp8ColBlockBytes;
}

// Development aid (tools/gemm_trace.cu compiles this TU with -DB200_GEMM_TRACE): per-role clock64 stamps of the
// first kTraceCtas CTAs.  Absent from the product build.
#ifdef B200_GEMM_TRACE
constexpr int kTraceCtas = 8, kTraceRoles = 8, kTraceSlots = 128;
__device__ long long g_trace[kTraceCtas][kTraceRoles][kTraceSlots];
#define GEMM_TRACE(role, slot)                                                                                         \
    do {                                                                                                               \
        const int _c = blockIdx.y * gridDim.x + blockIdx.x;                                                            \
        if (_c < kTraceCtas && (slot) < kTraceSlots) g_trace[_c][role][slot] = clock64();                              \
    } while (0)
// wall-clock (ns) begin / end of EVERY CTA of the last kWallLaunches launches: the anatomy of back-to-back kernels under
// a CUDA graph + PDL (how long after kernel i's last CTA ends does kernel i+1's first CTA start, how skewed are the starts)
constexpr int kWallLaunches = 16, kWallCtas = 160;
__device__ unsigned long long g_wall[kWallLaunches][kWallCtas][2];
B200_DEVICE unsigned long long wall_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GEMM_WALL(p, which)                                                                                            \
    do {                                                                                                               \
        const int _c = blockIdx.y * gridDim.x + blockIdx.x;                                                            \
        if (threadIdx.x == 0 && _c < kWallCtas) g_wall[(p).trace_launch % kWallLaunches][_c][which] = wall_ns();       \
    } while (0)
// watchdog wait: records (tag, cta, iteration) of the first wait that does not complete and traps
__device__ int           g_hang[8];
__device__ volatile int* g_hang_host;  // mapped pinned host memory: survives a dead context
B200_DEVICE void mbar_wait_dbg(uint32_t bar, uint32_t parity, int tag, int it)
{
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (*(volatile int*)&g_hang[0] != 0) {
            return;  // somebody already timed out: drain the kernel
        }
        if (clock64() - t0 > 200000000ll) {
            if (atomicCAS(&g_hang[0], 0, 1) == 0) {
                g_hang[1] = tag, g_hang[2] = blockIdx.x, g_hang[3] = blockIdx.y, g_hang[4] = it, g_hang[5] = threadIdx.x;
                g_hang[6] = (int)parity;
                if (g_hang_host) {
                    for (int q = 1; q < 7; ++q) g_hang_host[q] = g_hang[q];
                    g_hang_host[0] = 1;
                    __threadfence_system();
                }
                __threadfence();
            }
            return;
        }
    }
}
#define GEMM_WAIT(bar, parity, tag, it) mbar_wait_dbg(bar, parity, tag, it)
#else
#define GEMM_TRACE(role, slot)                                                                                         \
    do {                                                                                                               \
    } while (0)
#define GEMM_WALL(p, which)                                                                                            \
    do {                                                                                                               \
    } while (0)
#define GEMM_WAIT(bar, parity, tag, it) mbar_wait(bar, parity)
#endif

struct Params {
    const uint8_t* w;   // prepared weight stream
    const void*    xp;  // packed activations [K/8][MT][8] T
    void*          y;
    const void*    bias;      // [N] T or null
    float*         partials;  // split-K exchange [n-group][split][MT][128 * TILES] fp32 (L2 resident)
    int*           counters;  // [n-group][2] arrive / done, zero before first use; left zero by the kernel
    int            N, K;
    int            MT;  // padded token count (multiple of 16, <= 128)
    int            M;   // real tokens
    int            splits;  // K splits along grid.y (1 .. 16); > 1 needs a cooperative launch (co-residency)
    int            nw, nx;  // W / X ring depths
    int            na;      // TMEM A-ring slots (multiple of TILES)
    int            epilogue;  // 0 none, 1 gated SiLU (columns interleaved gate/up -> N/2 outputs),
                              // 2 residual add + RMSNorm of the full row (splits > 1 only; see epilogue part 2)
    void*          residual;  // epilogue 2: [M][N] T, updated in place (r += y)
    const void*    norm_w;    // epilogue 2: [N] T
    float          norm_eps;
    float*         ssq;       // epilogue 2: [n-group][MT] partial sums of squares (L2 exchange)
    int*           norm_ctr;  // epilogue 2: [splits][2] arrive / done counters of the n-groups owning the same tokens
    int            out_mode;  // 0 row-major [M][ldy]; 1 packed [Nout/8][out_mt][8]
    int            ldy;
    int            out_mt;
    const int*     u4_fast;    // u4 + f16 only: trailer word of the prepared stream, 1 = "subnormal" layout (see kU4FastShift)
    int            tmem_cols;  // 256 or 512 (power of two)
    int            a_col;      // first TMEM column of the A-operand ring
#ifdef B200_GEMM_TRACE
    int            trace_launch;  // tools/gemm_trace.cu: running launch number (GEMM_WALL)
#endif
#ifdef B200_GEMM_TOOL
    int            variant;    // tools/gemm_sweep.cu only: bit 0 = the W producer issues a slot's copies from several
                               // lanes at once, bit 1 = ... and each tile's block as two half-size copies,
                               // bit 2 = the X producer issues the activation tile as four copies from four lanes
#endif
};

B200_DEVICE int w_block_bytes(int fmt)
{
    return block_bytes_of(fmt);
}

// two e4m3 bytes (low 16 bits of `x`) -> two floats (exact): cvt.rn.f16x2.e4m3x2 then f16 -> f32
B200_DEVICE float2 cvt_e4m3x2_f32(uint32_t x)
{
    uint32_t h2;
    asm("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tcvt.rn.f16x2.e4m3x2 %0, t;\n\t}" : "=r"(h2) : "r"(x));
    const half2 h = reinterpret_cast<const half2&>(h2);
    return make_float2(__low2float(h), __high2float(h));
}

template<class T>
B200_DEVICE void store_out(const Params& p, int tok, int col, float v)
{
    using Ops = TypeOps<T>;
    T* y      = reinterpret_cast<T*>(p.y);
    if (p.out_mode == 0) {
        y[(size_t)tok * p.ldy + col] = Ops::from_f32(v);
    }
    else {
        y[((size_t)(col >> 3) * p.out_mt + tok) * 8 + (col & 7)] = Ops::from_f32(v);
    }
}

// one accumulator value of row n_g (all 32 lanes of the warp call this with the same tok)
template<class T>
B200_DEVICE void finalize(const Params& p, int tok, int n_g, float acc)
{
    using Ops = TypeOps<T>;
    if (p.bias) {
        acc += Ops::to_f32(reinterpret_cast<const T*>(p.bias)[n_g]);
    }
    if (p.epilogue == 1) {
        const float other = __shfl_xor_sync(0xffffffffu, acc, 1);
        if ((n_g & 1) == 0 && tok < p.M) {
            const float silu = __fdividef(acc, 1.f + __expf(-acc));
            store_out<T>(p, tok, n_g >> 1, silu * other);
        }
    }
    else if (tok < p.M) {
        store_out<T>(p, tok, n_g, acc);
    }
}

// two adjacent accumulator values (columns n_g, n_g + 1; n_g even) of one token
template<class T>
B200_DEVICE void finalize_pair(const Params& p, int tok, int n_g, float a0, float a1)
{
    using Ops = TypeOps<T>;
    if (tok >= p.M) {
        return;
    }
    if (p.bias) {
        a0 += Ops::to_f32(reinterpret_cast<const T*>(p.bias)[n_g]);
        a1 += Ops::to_f32(reinterpret_cast<const T*>(p.bias)[n_g + 1]);
    }
    if (p.epilogue == 1) {
        const float silu = __fdividef(a0, 1.f + __expf(-a0));
        store_out<T>(p, tok, n_g >> 1, silu * a1);
    }
    else {
        T* y = reinterpret_cast<T*>(p.y);
        T* d = p.out_mode == 0 ? y + (size_t)tok * p.ldy + n_g : y + ((size_t)(n_g >> 3) * p.out_mt + tok) * 8 + (n_g & 7);
        *reinterpret_cast<uint32_t*>(d) = Ops::pack2(a0, a1);
    }
}

B200_DEVICE void red_release_gpu_add(int* p, int v)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
B200_DEVICE void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
B200_DEVICE int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// TILES = n-tiles (128 output features each) per CTA.  TILES == 2 shares every activation tile between two weight
// tiles: the X stream (2x the W bytes at 64 tokens) is what saturates the L2->SM path, so this halves the ingest
// per weight byte.  TMEM: D accumulators of tile t at column t * d_stride, then an A ring of p.na 64-column slots
// used round-robin by (k-group, tile).
template<class T, int FMT, int TILES>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const Params p)
{
    using Ops = TypeOps<T>;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_wfull[kMaxW], bar_wempty[kMaxW];
    __shared__ uint64_t bar_xfull[kMaxX], bar_xempty[kMaxX];
    __shared__ uint64_t bar_aready[kASlots], bar_aempty[kASlots];
    __shared__ uint64_t bar_dready;
    __shared__ uint32_t s_tmem;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform
    const int nt0 = blockIdx.x * TILES, sp = blockIdx.y;
    const int kgs = p.K / kKG;
    const int kg0 = kgs * sp / p.splits, kg1 = kgs * (sp + 1) / p.splits;
    const int nk = kg1 - kg0;
    if (threadIdx.x == 64) GEMM_TRACE(4, 0);
    GEMM_WALL(p, 0);
    // PDL: let the next kernel start launching; everything up to the first activation load (barrier init, TMEM
    // allocation, and the whole weight prefetch -- weights are constants) runs before griddepcontrol.wait.
    pdl_launch_dependents();

    constexpr int  WB       = block_bytes_of(FMT);
    constexpr int  WB_pad   = (WB + 1023) & ~1023;
    constexpr int  WSLOT    = WB_pad * TILES;
    constexpr bool kDequant = FMT != kDense;
    constexpr int  kMmaWarps = TILES >= 2 ? 2 : 1;  // MMA-issuing warps (TILES == 4, tool builds: two tiles per warp)
    static_assert(TILES == 1 || TILES == 2 || TILES == 4, "1, 2 or (tool builds) 4 n-tiles per CTA");
    const int      NAt      = p.na / TILES;  // A-ring slots per tile (tile t owns slots [t * NAt, (t + 1) * NAt))
    static_assert(kDequant || TILES == 1, "dense weights use one tile per CTA");
    const int      XB       = p.MT * 256;  // 16 k8-chunks x MT x 16 B
    const uint32_t w_ring   = smem_u32(smem);
    const uint32_t x_ring   = w_ring + (uint32_t)p.nw * WSLOT;
    const int      d_stride = p.MT <= 64 ? 64 : 128;

    if (threadIdx.x == 64) GEMM_TRACE(7, 4);
    // one barrier per thread (a single thread initialising all 37 costs ~800 cycles before the first weight copy can be issued)
    {
        const int t = threadIdx.x;
        if (t < kMaxW) {
            mbar_init(smem_u32(&bar_wfull[t]), 1);
        }
        else if (t < 2 * kMaxW) {
            mbar_init(smem_u32(&bar_wempty[t - kMaxW]), kDequant ? kDqWarps * 32 : 1);
        }
        else if (t < 2 * kMaxW + kMaxX) {
            mbar_init(smem_u32(&bar_xfull[t - 2 * kMaxW]), 1);
        }
        else if (t < 2 * kMaxW + 2 * kMaxX) {
            mbar_init(smem_u32(&bar_xempty[t - 2 * kMaxW - kMaxX]), kMmaWarps);
        }
        else if (t < 2 * kMaxW + 2 * kMaxX + kASlots) {
            mbar_init(smem_u32(&bar_aready[t - 2 * kMaxW - 2 * kMaxX]), kDqWarps * 32 / TILES);
        }
        else if (t < 2 * kMaxW + 2 * kMaxX + 2 * kASlots) {
            mbar_init(smem_u32(&bar_aempty[t - 2 * kMaxW - 2 * kMaxX - kASlots]), 1);
        }
        else if (t == 2 * kMaxW + 2 * kMaxX + 2 * kASlots) {
            mbar_init(smem_u32(&bar_dready), kMmaWarps);
        }
        if (t <= 2 * kMaxW + 2 * kMaxX + 2 * kASlots) {
            fence_mbar_init();
        }
    }
    __syncthreads();

    if (warp == 0) {
        // ================================ W producer (starts before TMEM is allocated) ================================
#ifdef B200_GEMM_TOOL
        if (p.variant & 3) {
            // Experiment: is the producer bound by the ~300 cycles it takes ONE lane to issue a bulk copy?  Lane 0 waits
            // for the slot and arms the barrier; the copies of the slot are then issued by 2 or 4 lanes in parallel.
            const uint64_t pol_w  = make_evict_first_policy();
            const int      parts  = (p.variant & 2) ? 2 : 1;         // copies per tile block
            const int      PB     = WB / parts;                      // 8704 / 2 = 4352: still a multiple of 16
            const int      n_copy = TILES * parts;
            int            s      = 0;
            uint32_t       ph     = 0;
            const uint8_t* src    = p.w + ((size_t)nt0 * kgs + kg0) * WB;
            for (int i = 0; i < nk; ++i) {
                const uint32_t bar = smem_u32(&bar_wfull[s]);
                if (lane == 0) {
                    GEMM_WAIT(smem_u32(&bar_wempty[s]), ph ^ 1, 1, i);
                    mbar_expect_tx(bar, WB * TILES);
                }
                __syncwarp();
                if (lane < n_copy) {
                    const int t = lane / parts, h = lane % parts;
                    bulk_g2s_hint(w_ring + (uint32_t)s * WSLOT + t * WB_pad + h * PB, src + ((size_t)t * kgs + i) * WB + h * PB,
                                  PB, bar, pol_w);
                }
                if (++s == p.nw) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        else
#endif
        if (lane == 0) {
            const uint64_t pol_w = make_evict_first_policy();
            int      s  = 0;
            uint32_t ph = 0;
            const uint8_t* src = p.w + ((size_t)nt0 * kgs + kg0) * WB;
            for (int i = 0; i < nk; ++i) {
                GEMM_WAIT(smem_u32(&bar_wempty[s]), ph ^ 1, 1, i);
                GEMM_TRACE(0, i);
                const uint32_t bar = smem_u32(&bar_wfull[s]);
                mbar_expect_tx(bar, WB * TILES);
#pragma unroll
                for (int t = 0; t < TILES; ++t) {
                    bulk_g2s_hint(w_ring + (uint32_t)s * WSLOT + t * WB_pad, src + ((size_t)t * kgs + i) * WB, WB, bar,
                                  pol_w);
                }
                if (++s == p.nw) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    }
    else if (warp == kXWarp) {
        // ================================ X producer ================================
#ifdef B200_GEMM_TOOL
        if (p.variant & 4) {
            // Experiment: the activation tile of a slot as four quarter-size copies issued by four lanes.
            int            s   = 0;
            uint32_t       ph  = 0;
            const int      QB  = XB / 4;  // MT * 64 bytes: a multiple of 16
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.xp) + (size_t)kg0 * XB;
            pdl_wait();
            for (int i = 0; i < nk; ++i) {
                const uint32_t bar = smem_u32(&bar_xfull[s]);
                if (lane == 0) {
                    GEMM_WAIT(smem_u32(&bar_xempty[s]), ph ^ 1, 6, i);
                    mbar_expect_tx(bar, XB);
                }
                __syncwarp();
                if (lane < 4) {
                    bulk_g2s(x_ring + (uint32_t)s * XB + lane * QB, src + (size_t)i * XB + lane * QB, QB, bar);
                }
                if (++s == p.nx) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        else
#endif
        if (lane == 0) {
            int      s  = 0;
            uint32_t ph = 0;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.xp) + (size_t)kg0 * XB;
            pdl_wait();  // the activations come from the previous kernel
            for (int i = 0; i < nk; ++i) {
                GEMM_WAIT(smem_u32(&bar_xempty[s]), ph ^ 1, 6, i);
                GEMM_TRACE(6, i);
                const uint32_t bar = smem_u32(&bar_xfull[s]);
                mbar_expect_tx(bar, XB);
                bulk_g2s(x_ring + (uint32_t)s * XB, src + (size_t)i * XB, XB, bar);
                if (++s == p.nx) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    }
    else {
        // the MMA and dequant warps need the TMEM base address
        if (warp == 1) {
            umma::tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
        }
        umma::fence_before_sync();
        named_bar_sync(1, (2 + kDqWarps) * 32);
        umma::fence_after_sync();
        const uint32_t tmem = s_tmem;
        if (threadIdx.x == 64) GEMM_TRACE(4, 1);

        if (warp == 1 || warp == kMmaWarp2) {
            // ================================ MMA issuers (one warp per n-tile) ================================
            // The whole warp runs this loop under warp-uniform control flow and one elected lane issues; every
            // operand is derived from uniform values, so ptxas keeps descriptors in uniform registers and a k-step
            // costs two uniform adds + UTCHMMA.  A UTCHMMA still costs ~50 issue cycles, more than the 32 tensor
            // cycles of a 128x64x16 MMA, so each tile of a two-tile CTA gets its own issuing warp.
            const int t = warp == 1 ? 0 : 1;
            if constexpr (TILES == 4) {
                // Four tiles share each activation tile (ingest 1.5 W instead of 2 W per weight byte).  Each issuing warp
                // owns two tiles and alternates between them inside a k-group; TMEM leaves one A slot per tile
                // (4 x 64 accumulator columns + 4 x 64), so a tile's next de-quantisation overlaps the other tile's MMAs.
                const uint32_t idesc  = umma::make_idesc_f16(kBN, p.MT, std::is_same<T, nv_bfloat16>::value);
                const uint32_t b_step = (uint32_t)(2 * p.MT * 16) >> 4;
                int            sx = 0, sa = 0;
                uint32_t       phx = 0, pha = 0;
                for (int i = 0; i < nk; ++i) {
                    GEMM_WAIT(smem_u32(&bar_xfull[sx]), phx, 2, i);
                    const uint64_t bd0 = umma::make_smem_desc(x_ring + (uint32_t)sx * XB, p.MT * 16, 128);
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt) {
                        const int      tile = 2 * t + tt;
                        const int      slot = tile * NAt + sa;
                        const uint32_t d_tm = tmem + tile * d_stride;
                        const uint32_t a_tm = tmem + p.a_col + slot * 64;
                        GEMM_WAIT(smem_u32(&bar_aready[slot]), pha, 3, i);
                        umma::fence_after_sync();
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < kKG / 16; ++ks) {
                                umma::mma_f16_ts(d_tm, a_tm + ks * 8, bd0 + (uint64_t)(ks * b_step), idesc,
                                                 (i > 0 || ks > 0) ? 1u : 0u);
                            }
                            umma::commit(smem_u32(&bar_aempty[slot]));
                            if (tt == 1) {
                                umma::commit(smem_u32(&bar_xempty[sx]));
                            }
                        }
                        __syncwarp();
                    }
                    if (++sa == NAt) {
                        sa = 0;
                        pha ^= 1;
                    }
                    if (++sx == p.nx) {
                        sx = 0;
                        phx ^= 1;
                    }
                }
                if (elect_one()) {
                    umma::commit(smem_u32(&bar_dready));
                }
                __syncwarp();
            }
            else if (t < TILES) {
                const uint32_t idesc  = umma::make_idesc_f16(kBN, p.MT, std::is_same<T, nv_bfloat16>::value);
                const uint32_t b_step = (uint32_t)(2 * p.MT * 16) >> 4;  // descriptor start-address units per k-step
                const uint32_t a_step = (uint32_t)(2 * kBN * 16) >> 4;
                const uint32_t d_tm   = tmem + t * d_stride;
                const int      a0     = t * NAt;
                int            sw = 0, sx = 0, sa = 0;
                uint32_t       phw = 0, phx = 0, pha = 0;
                for (int i = 0; i < nk; ++i) {
                    GEMM_WAIT(smem_u32(&bar_xfull[sx]), phx, 2, i);
                    if (kDequant) {
                        GEMM_WAIT(smem_u32(&bar_aready[a0 + sa]), pha, 3, i);
                    }
                    else {
                        GEMM_WAIT(smem_u32(&bar_wfull[sw]), phw, 7, i);
                    }
                    umma::fence_after_sync();
                    if (lane == 0 && t == 0) GEMM_TRACE(3, i);
                    const uint64_t bd0  = umma::make_smem_desc(x_ring + (uint32_t)sx * XB, p.MT * 16, 128);
                    const uint64_t ad0  = umma::make_smem_desc(w_ring + (uint32_t)sw * WSLOT, kBN * 16, 128);
                    const uint32_t a_tm = tmem + p.a_col + (a0 + sa) * 64;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < kKG / 16; ++ks) {
                            const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                            if (kDequant) {
                                umma::mma_f16_ts(d_tm, a_tm + ks * 8, bd0 + (uint64_t)(ks * b_step), idesc, acc);
                            }
                            else {
                                umma::mma_f16_ss(d_tm, ad0 + (uint64_t)(ks * a_step), bd0 + (uint64_t)(ks * b_step),
                                                 idesc, acc);
                            }
                        }
                        if (kDequant) {
                            umma::commit(smem_u32(&bar_aempty[a0 + sa]));
                        }
                        else {
                            umma::commit(smem_u32(&bar_wempty[sw]));
                        }
                        umma::commit(smem_u32(&bar_xempty[sx]));
                    }
                    __syncwarp();
                    if (lane == 0 && t == 0) GEMM_TRACE(5, i);
                    if (++sa == NAt) {
                        sa = 0;
                        pha ^= 1;
                    }
                    if (++sx == p.nx) {
                        sx = 0;
                        phx ^= 1;
                    }
                    if (++sw == p.nw) {
                        sw = 0;
                        phw ^= 1;
                    }
                }
                if (elect_one()) {
                    umma::commit(smem_u32(&bar_dready));
                }
                __syncwarp();
            }
        }
        else {
            // ================================ dequant warps ================================
            // TILES == 1: 16 warps share the tile (4 per TMEM lane quarter, one 32-wide k-chunk each).
            // TILES == 2: warps 2-9 own tile 0, warps 10-17 tile 1 (2 per lane quarter, two k-chunks each), so the
            // two tiles of a k-group are de-quantised concurrently and every thread still observes every phase of the
            // barriers it waits on.
            const int      dw        = warp - 2;
            // TILES == 4 (tool builds): 4 warps per tile (one per lane quarter), all four k-chunks each.
            const int      my_tile   = TILES == 4 ? (dw >> 2) : TILES == 2 ? (dw >> 3) : 0;
            const int      kc0       = TILES == 4 ? 0 : TILES == 2 ? ((dw >> 2) & 1) * 2 : (dw >> 2);  // first 32-wide k-chunk
            constexpr int  NKC       = TILES == 4 ? 4 : TILES == 2 ? 2 : 1;                            // k-chunks per thread
            const int      lane_base = 32 * (warp & 3);   // TMEM lanes this warp may touch
            const int      n_local   = lane_base + lane;
            const uint32_t trow      = tmem + ((uint32_t)lane_base << 16);
            if constexpr (kDequant) {
                const int a0 = my_tile * NAt;
                int       sw = 0, sa = 0;
                uint32_t  phw = 0, pha = 0;
                // software pipeline: the packed weights of k-group i + 1 are fetched from shared memory while the
                // tcgen05.st of k-group i is still in flight (one thread = one weight row x NKC 32-wide k-chunks)
                constexpr int UNITS = FMT == kU4 ? 1 : 2;  // 16-byte units per 32-wide k-chunk (u4: 32 nibbles; fp8: 2 x 16 B)
                uint32_t      st2   = 0;                   // u4: (scale, -zero*scale) as T2;  fp8: block scale as T (low half)
                uint4         v[NKC][UNITS];
                auto          fetch = [&](int i, int slot, uint32_t phase) {
                    GEMM_WAIT(smem_u32(&bar_wfull[slot]), phase, 4, i);
                    if (warp == 2 && lane == 0) GEMM_TRACE(1, i);
                    const uint32_t wsm = w_ring + (uint32_t)slot * WSLOT + my_tile * WB_pad;
                    if constexpr (FMT == kU4) {
                        st2 = lds32(wsm + 8192 + n_local * 4);
#pragma unroll
                        for (int q = 0; q < NKC; ++q) v[q][0] = lds128(wsm + ((kc0 + q) * 128 + n_local) * 16);
                    }
                    else {
                        st2 = lds32(wsm + 16384 + (FMT == kFp8Col ? n_local * 4 : 0));
#pragma unroll
                        for (int q = 0; q < NKC; ++q) {
                            v[q][0] = lds128(wsm + ((2 * (kc0 + q)) * 128 + n_local) * 16);
                            v[q][1] = lds128(wsm + ((2 * (kc0 + q) + 1) * 128 + n_local) * 16);
                        }
                    }
                };
                // One 32-wide k-chunk of this thread's row: packed weights -> 16 registers of T pairs
                auto convert = [&](auto DN, int q, uint32_t s2, uint32_t t2, uint32_t (&r)[16]) {
                    if constexpr (FMT == kU4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            uint32_t h[4];
                            if constexpr (decltype(DN)::value) {
                                const uint32_t x = (&v[q][0].x)[u];
                                h[0] = (x << 4) & kU4FastMask;
                                h[1] = x & kU4FastMask;
                                h[2] = (x >> 4) & kU4FastMask;
                                h[3] = __funnelshift_r(x, x, 8) & kU4FastMask;
                            }
                            else {
                                Ops::cvt_u4x8((&v[q][0].x)[u], h);
                            }
#pragma unroll
                            for (int x = 0; x < 4; ++x) r[u * 4 + x] = Ops::fma2(h[x], s2, t2);
                        }
                    }
                    else {
                        // w = T(e4m3) *_T T(block scale): both factors are exact in fp32, so one fp32 multiply
                        // rounded once to T is the T multiply of the reference (transform.h:92-97)
                        const float sc = Ops::to_f32(Ops::from_bits((uint16_t)(st2 & 0xffffu)));
#pragma unroll
                        for (int u = 0; u < UNITS; ++u) {
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const uint32_t w4 = (&v[q][u].x)[x];
                                const float2   lo = cvt_e4m3x2_f32(w4 & 0xffffu);
                                const float2   hi = cvt_e4m3x2_f32(w4 >> 16);
                                r[u * 8 + x * 2]     = Ops::pack2(lo.x * sc, lo.y * sc);
                                r[u * 8 + x * 2 + 1] = Ops::pack2(hi.x * sc, hi.y * sc);
                            }
                        }
                    }
                };
                // layout flag of the prepared stream (a constant of the weight: read before griddepcontrol.wait matters not)
                bool dn = false;
                if constexpr (FMT == kU4 && std::is_same<T, half>::value) {
                    dn = p.u4_fast != nullptr && __ldg(p.u4_fast) != 0;
                }
                if (nk > 0) {
                    fetch(0, 0, 0);
                }
                // Fast path (u4, W ring of exactly kFastW slots, kASlots A slots): the loop is unrolled over one lap of the W
                // ring, so every ring position -- barrier addresses, shared-memory and TMEM offsets, the A ring's phase --
                // is a compile-time constant.  ncu (r02_gemm.ncu-rep, w1w3): the schedulers of an active SM issue 96 % of
                // the cycles and 55 of the 162 instructions a de-quantisation warp spends per k-group were ring
                // bookkeeping (uniform-datapath address arithmetic, wrap / phase selects) re-derived every group.
                constexpr int kFastW = 8;
                constexpr int NAT_C  = kASlots / TILES;
                if (FMT == kU4 && TILES <= 2 && p.nw == kFastW && NAt == NAT_C) {
                    const uint32_t aempty0 = smem_u32(&bar_aempty[0]) + a0 * 8, aready0 = smem_u32(&bar_aready[0]) + a0 * 8;
                    const uint32_t wempty0 = smem_u32(&bar_wempty[0]);
                    const uint32_t tm_a    = trow + p.a_col + a0 * 64 + kc0 * 16;
                    auto           lap_loop = [&](auto DN) {
                        uint32_t lap = 0;  // parity of the W ring's lap = (i / kFastW) & 1
                        for (int i0 = 0; i0 < nk; i0 += kFastW) {
#pragma unroll
                            for (int U = 0; U < kFastW; ++U) {
                                const int i = i0 + U;
                                if (i >= nk) {
                                    break;
                                }
                                const int      SA = U % NAT_C;               // compile-time after unrolling
                                const uint32_t pa = ((U / NAT_C) & 1) ^ 1u;  // kFastW / NAT_C is even: the A phase restarts every lap
                                const uint32_t s2 = __byte_perm(st2, st2, 0x1010), t2 = __byte_perm(st2, st2, 0x3232);
                                GEMM_WAIT(aempty0 + SA * 8, pa, 8, i);
                                umma::fence_after_sync();
#pragma unroll
                                for (int q = 0; q < NKC; ++q) {
                                    uint32_t r[16];
                                    convert(DN, q, s2, t2, r);
                                    umma::st_32x32b_x16(tm_a + SA * 64 + q * 16, r);
                                }
                                mbar_arrive(wempty0 + U * 8);
                                if (i + 1 < nk) {
                                    fetch(i + 1, (U + 1) % kFastW, U + 1 == kFastW ? lap ^ 1u : lap);
                                }
                                umma::wait_st();
                                umma::fence_before_sync();
                                mbar_arrive(aready0 + SA * 8);
                                if (warp == 2 && lane == 0) GEMM_TRACE(2, i);
                            }
                            lap ^= 1u;
                        }
                    };
                    if (dn) {
                        lap_loop(std::true_type{});
                    }
                    else {
                        lap_loop(std::false_type{});
                    }
                }
                else {
                    auto ring_loop = [&](auto DN) {
                        for (int i = 0; i < nk; ++i) {
                            const uint32_t s2 = __byte_perm(st2, st2, 0x1010), t2 = __byte_perm(st2, st2, 0x3232);
                            // the TMEM slot must have been drained by the MMAs that used it last
                            GEMM_WAIT(smem_u32(&bar_aempty[a0 + sa]), pha ^ 1, 8, i);
                            umma::fence_after_sync();
#pragma unroll
                            for (int q = 0; q < NKC; ++q) {
                                uint32_t r[16];
                                convert(DN, q, s2, t2, r);
                                umma::st_32x32b_x16(trow + p.a_col + (a0 + sa) * 64 + (kc0 + q) * 16, r);
                            }
                            // the W slot is free once its bytes are in registers (the arrive releases the reads)
                            mbar_arrive(smem_u32(&bar_wempty[sw]));
                            if (++sw == p.nw) {
                                sw = 0;
                                phw ^= 1;
                            }
                            if (i + 1 < nk) {
                                fetch(i + 1, sw, phw);
                            }
                            umma::wait_st();
                            umma::fence_before_sync();
                            mbar_arrive(smem_u32(&bar_aready[a0 + sa]));
                            if (warp == 2 && lane == 0) GEMM_TRACE(2, i);
                            if (++sa == NAt) {
                                sa = 0;
                                pha ^= 1;
                            }
                        }
                    };
                    if (dn) {
                        ring_loop(std::true_type{});
                    }
                    else {
                        ring_loop(std::false_type{});
                    }
                }
            }
            // epilogue role: (tile, token half) of the accumulator rows in this warp's lane quarter
            const int  epi_tile   = my_tile;
            const int  epi_half   = TILES == 4 ? 0 : TILES == 2 ? ((dw >> 2) & 1) : (dw >> 2);
            const bool epi_active = TILES >= 2 || (dw >> 2) < 2;
            // ============================== epilogue, part 1 ==============================
            // splits == 1: straight to the output.  splits > 1: fp32 partials to the L2-resident exchange buffer.
            GEMM_WAIT(smem_u32(&bar_dready), 0, 5, 0);
            umma::fence_after_sync();
            pdl_wait();  // already satisfied (the MMAs consumed upstream activations); orders this thread's global
                         // accesses (output / exchange buffers may still be read by an upstream kernel otherwise)
            if (threadIdx.x == 64) GEMM_TRACE(4, 2);
            const int half = TILES == 4 ? p.MT : p.MT >> 1;  // TILES == 4: one warp per (tile, lane quarter) takes every token
            const int tb = epi_half * half, te = tb + half;
            const uint32_t drow = trow + epi_tile * d_stride;
            if (p.splits == 1) {
                // Accumulators -> fp32 staging tile [token][row] in the (now idle) operand rings: thread = row, so a warp's 32
                // stores of one token are consecutive words (conflict-free); the direct epilogue below then works token-major
                // with 128-bit accesses (measured on w1w3: 3.6 us -> 0.5 us for this phase).  Row pitch ROWS + 4 words: the
                // token-fastest 128-bit reads of a quarter-warp hit distinct bank groups.
                if (epi_active) {
                    constexpr int  ROWS  = kBN * TILES;
                    const int      pitch = ROWS + 4;
                    const uint32_t dst   = w_ring + (uint32_t)(epi_tile * kBN + n_local) * 4;
                    for (int t0 = tb; t0 < te; t0 += 32) {
                        uint32_t v[2][16];
                        umma::ld_32x32b_x16(drow + t0, v[0]);
                        if (t0 + 16 < te) {
                            umma::ld_32x32b_x16(drow + t0 + 16, v[1]);
                        }
                        umma::wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (t0 + i < te) {
                                sts32(dst + (uint32_t)((t0 + i) * pitch) * 4, v[i >> 4][i & 15]);
                            }
                        }
                    }
                }
            }
            else if (epi_active) {
                // fp32 partial tile -> global exchange buffer [group][split][tok][row] (stays in L2), 128-byte rows per warp
                // store.  (Staging the tile in shared memory and moving it with ONE bulk store / bulk loads was measured:
                // 3.7k cycles for the 64 KB store alone against 2.2k for these direct stores -- the TMA engine moves a single
                // large copy at ~17 B/clk.)
                constexpr int ROWS = kBN * TILES;
                float* part = p.partials + ((size_t)blockIdx.x * p.splits + sp) * p.MT * ROWS + epi_tile * kBN + n_local;
                for (int t0 = tb; t0 < te; t0 += 32) {
                    uint32_t v[2][16];
                    umma::ld_32x32b_x16(drow + t0, v[0]);
                    if (t0 + 16 < te) {
                        umma::ld_32x32b_x16(drow + t0 + 16, v[1]);
                    }
                    umma::wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (t0 + i < te) {
                            __stcg(part + (size_t)(t0 + i) * ROWS, __uint_as_float(v[i >> 4][i & 15]));
                        }
                    }
                }
            }
            if (threadIdx.x == 64) GEMM_TRACE(4, 3);
        }
    }
    if (p.splits == 1 && warp >= 2 && warp < kXWarp) {
        // ============================== epilogue, part 2 (no split): staging tile -> bias / gated SiLU -> global ==============================
        // unit = 8 output columns of one token: 128-bit reads of the staging tile, one 16-byte store
        named_bar_sync(2, kDqWarps * 32);
        constexpr int ROWS  = kBN * TILES;
        const int     pitch = ROWS + 4;
        const bool    gated = p.epilogue == 1;
        const int     OC    = gated ? ROWS / 2 : ROWS;                    // output columns of this CTA
        const int     c0    = gated ? nt0 * (kBN / 2) : nt0 * kBN;        // its first output column
        const int     tid   = threadIdx.x - 64;
        const int     UC    = OC / 8;                                     // units per token
        T*            y     = reinterpret_cast<T*>(p.y);
        const T*      bias  = reinterpret_cast<const T*>(p.bias);
        for (int u = tid; u < UC * p.M; u += kDqWarps * 32) {
            // packed output: tokens fastest (a warp writes 512 contiguous bytes); row-major: columns fastest
            const int c   = p.out_mode == 1 ? u / p.M : u % UC;
            const int tok = p.out_mode == 1 ? u - c * p.M : u / UC;
            float     o[8];
            if (gated) {
                const uint32_t src = w_ring + (uint32_t)(tok * pitch + c * 16) * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 v = lds128(src + q * 16);
                    float       g0 = __uint_as_float(v.x), u0 = __uint_as_float(v.y);
                    float       g1 = __uint_as_float(v.z), u1 = __uint_as_float(v.w);
                    if (bias) {
                        const int n = nt0 * kBN + c * 16 + q * 4;
                        g0 += Ops::to_f32(bias[n]), u0 += Ops::to_f32(bias[n + 1]);
                        g1 += Ops::to_f32(bias[n + 2]), u1 += Ops::to_f32(bias[n + 3]);
                    }
                    o[2 * q]     = __fdividef(g0, 1.f + __expf(-g0)) * u0;
                    o[2 * q + 1] = __fdividef(g1, 1.f + __expf(-g1)) * u1;
                }
            }
            else {
                const uint32_t src = w_ring + (uint32_t)(tok * pitch + c * 8) * 4;
                const uint4    a = lds128(src), b2 = lds128(src + 16);
                o[0] = __uint_as_float(a.x), o[1] = __uint_as_float(a.y), o[2] = __uint_as_float(a.z), o[3] = __uint_as_float(a.w);
                o[4] = __uint_as_float(b2.x), o[5] = __uint_as_float(b2.y), o[6] = __uint_as_float(b2.z), o[7] = __uint_as_float(b2.w);
                if (bias) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] += Ops::to_f32(bias[nt0 * kBN + c * 8 + q]);
                }
            }
            uint4 pk;
            pk.x = Ops::pack2(o[0], o[1]), pk.y = Ops::pack2(o[2], o[3]), pk.z = Ops::pack2(o[4], o[5]), pk.w = Ops::pack2(o[6], o[7]);
            T* d = p.out_mode == 1 ? y + ((size_t)(c0 / 8 + c) * p.out_mt + tok) * 8 : y + (size_t)tok * p.ldy + c0 + c * 8;
            *reinterpret_cast<uint4*>(d) = pk;
        }
    }
    if (p.splits > 1 && warp >= 2 && warp < kXWarp) {
        // ============================== epilogue, part 2: split-K reduce-scatter through L2 ==============================
        // The `splits` CTAs of an n-group are co-resident (cooperative launch).  Each publishes its partial tile,
        // then sums ITS token slice over all partials in rank order (deterministic) and writes the output.
        constexpr int ROWS = kBN * TILES;
        int*          ctr  = p.counters + 2 * blockIdx.x;
        // Publish: the partial stores of all 512 threads are ordered before thread 64's gpu-scope RELEASE by the CTA barrier
        // (cumulativity), so no per-thread __threadfence (= fence.sc.gpu: measured ~0.9 us here) is needed; the readers
        // are ordered behind thread 64's ACQUIRE spin by the second barrier.
        named_bar_sync(2, kDqWarps * 32);
        const int t_lo = p.MT * sp / p.splits, t_hi = p.MT * (sp + 1) / p.splits;
        if (threadIdx.x == 64) {
            GEMM_TRACE(7, 0);
            red_release_gpu_add(ctr, 1);
            while (ld_acquire_gpu(ctr) < p.splits) {
            }
            GEMM_TRACE(7, 1);
        }
        named_bar_sync(2, kDqWarps * 32);
        // token ownership: split r sums tokens [MT * r / splits, MT * (r + 1) / splits)  (any split count works)
        const int    tid  = threadIdx.x - 64;  // 0 .. 32 * kDqWarps - 1
        const float* base = p.partials + (size_t)blockIdx.x * p.splits * p.MT * ROWS + (size_t)t_lo * ROWS;
        if (p.epilogue != 2) {
            for (int idx = tid; idx < (t_hi - t_lo) * (ROWS / 2); idx += kDqWarps * 32) {
                const int tl = idx / (ROWS / 2), np = idx - tl * (ROWS / 2);
                float2    acc = make_float2(0.f, 0.f);
                const float2* src = reinterpret_cast<const float2*>(base + tl * ROWS) + np;
                // eight partials in flight per batch; summed in split order: deterministic
                for (int r0 = 0; r0 < p.splits; r0 += 8) {
                    float2 v[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (r0 + r < p.splits) {
                            v[r] = __ldcg(src + (size_t)(r0 + r) * (p.MT * ROWS / 2));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (r0 + r < p.splits) {
                            acc.x += v[r].x;
                            acc.y += v[r].y;
                        }
                    }
                }
                finalize_pair<T>(p, t_lo + tl, nt0 * kBN + 2 * np, acc.x, acc.y);
            }
        }
        else {
            // ---- fused residual add + RMSNorm (replaces invokeBiasResidualRMSNorm, rms_norm.cu:287-363) ----
            // This CTA owns tokens [t_lo, t_hi) x ROWS columns; the row's sum of squares is completed through a
            // tiny L2 exchange with the other n-groups that own the same tokens (all co-resident).
            constexpr int kMaxItems = kNormMaxItems;  // host guarantees ntok * ROWS / 2 <= kMaxItems * 512
            constexpr int WPT       = ROWS / 64;      // warps per token (ROWS / 2 pairs, 32 per warp)
            __shared__ float s_ss[32][WPT];
            __shared__ float s_inv[32];
            const int ntok = t_hi - t_lo;
            const int wq   = ((threadIdx.x - 64) >> 5) % WPT;  // warp slot within a token
            float2    rr[kMaxItems];
            T*        res = reinterpret_cast<T*>(p.residual);
#pragma unroll
            for (int k = 0; k < kMaxItems; ++k) {
                const int idx = tid + k * (kDqWarps * 32);
                rr[k]         = make_float2(0.f, 0.f);
                if (idx < ntok * (ROWS / 2)) {
                    const int tl = idx / (ROWS / 2), np = idx - tl * (ROWS / 2);
                    const int tok = t_lo + tl, n_g = nt0 * kBN + 2 * np;
                    float2    acc = make_float2(0.f, 0.f);
                    for (int r = 0; r < p.splits; ++r) {
                        const float2 v =
                            __ldcg(reinterpret_cast<const float2*>(base + (size_t)r * p.MT * ROWS + tl * ROWS) + np);
                        acc.x += v.x;
                        acc.y += v.y;
                    }
                    if (p.bias) {
                        acc.x += Ops::to_f32(reinterpret_cast<const T*>(p.bias)[n_g]);
                        acc.y += Ops::to_f32(reinterpret_cast<const T*>(p.bias)[n_g + 1]);
                    }
                    float ss = 0.f;
                    if (tok < p.M) {
                        T h0 = Ops::from_f32(acc.x), h1 = Ops::from_f32(acc.y);  // the GEMM output, rounded to T
                        T* rp = res + (size_t)tok * p.N + n_g;
                        T  r2[2];
                        *reinterpret_cast<uint32_t*>(&r2[0]) = *reinterpret_cast<const uint32_t*>(rp);
                        r2[0] = Ops::add(r2[0], h0);  // residual add in T (rms_norm.cu:317)
                        r2[1] = Ops::add(r2[1], h1);
                        *reinterpret_cast<uint32_t*>(rp) = *reinterpret_cast<const uint32_t*>(&r2[0]);
                        rr[k] = make_float2(Ops::to_f32(r2[0]), Ops::to_f32(r2[1]));
                        ss    = fmaf(rr[k].x, rr[k].x, rr[k].y * rr[k].y);
                    }
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
                    if (lane == 0) {
                        s_ss[tl][wq] = ss;  // the 32 idx of a warp belong to one token (ROWS / 2 is a multiple of 32)
                    }
                }
            }
            named_bar_sync(2, kDqWarps * 32);
            if (tid < ntok) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < WPT; ++w) t += s_ss[tid][w];
                __stcg(p.ssq + (size_t)blockIdx.x * p.MT + t_lo + tid, t);
            }
            __threadfence();
            named_bar_sync(2, kDqWarps * 32);
            int* nctr = p.norm_ctr + 2 * sp;
            if (threadIdx.x == 64) {
                atomicAdd(nctr, 1);
                while (ld_acquire_gpu(nctr) < (int)gridDim.x) {
                }
            }
            named_bar_sync(2, kDqWarps * 32);
            if (tid < ntok) {
                float t = 0.f;
                for (int g = 0; g < (int)gridDim.x; ++g) t += __ldcg(p.ssq + (size_t)g * p.MT + t_lo + tid);  // fixed order
                s_inv[tid] = rsqrtf(t / (float)p.N + p.norm_eps);
            }
            named_bar_sync(2, kDqWarps * 32);
            const T* nw = reinterpret_cast<const T*>(p.norm_w);
            T*       y  = reinterpret_cast<T*>(p.y);
#pragma unroll
            for (int k = 0; k < kMaxItems; ++k) {
                const int idx = tid + k * (kDqWarps * 32);
                if (idx < ntok * (ROWS / 2)) {
                    const int tl = idx / (ROWS / 2), np = idx - tl * (ROWS / 2);
                    const int tok = t_lo + tl, n_g = nt0 * kBN + 2 * np;
                    if (tok < p.M) {
                        const float inv = s_inv[tl];
                        T           w2[2], o2[2];
                        *reinterpret_cast<uint32_t*>(&w2[0]) = *reinterpret_cast<const uint32_t*>(nw + n_g);
                        o2[0] = Ops::mul(Ops::from_f32(rr[k].x * inv), w2[0]);  // T(float(r) * inv_rms) * w
                        o2[1] = Ops::mul(Ops::from_f32(rr[k].y * inv), w2[1]);
                        T* d = p.out_mode == 0 ? y + (size_t)tok * p.ldy + n_g :
                                                 y + ((size_t)(n_g >> 3) * p.out_mt + tok) * 8 + (n_g & 7);
                        *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<const uint32_t*>(&o2[0]);
                    }
                }
            }
            named_bar_sync(2, kDqWarps * 32);
            if (threadIdx.x == 64) {
                if (atomicAdd(nctr + 1, 1) == (int)gridDim.x - 1) {
                    nctr[1] = 0;
                    __threadfence();
                    nctr[0] = 0;
                }
            }
        }
        named_bar_sync(2, kDqWarps * 32);
        if (threadIdx.x == 64) {
            GEMM_TRACE(7, 2);
            // the last CTA of the group to finish reading re-arms the counters for the next launch
            if (atomicAdd(ctr + 1, 1) == p.splits - 1) {
                ctr[1] = 0;
                __threadfence();
                ctr[0] = 0;
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        umma::tmem_dealloc(s_tmem, p.tmem_cols);
    }
    if (threadIdx.x == 64) GEMM_TRACE(7, 3);
    GEMM_WALL(p, 1);
}

}  // namespace gemm
}  // namespace b200
