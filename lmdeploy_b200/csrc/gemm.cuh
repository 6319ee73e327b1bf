// Weight-only-quant GEMM for decode (M = tokens <= 128) on tcgen05 tensor cores, sm_100a.
//
// Replaces, behind the C ABI, LlamaLinear::Forward -> gemm::Gemm::Run
//   (src/turbomind/models/llama/LlamaLinear.cu:140-216, kernels/gemm/gemm.cu:257-344) together with the
//   paired weight re-packer LinearWeight::prepare (models/linear_weight.cc:101-324).
// It is NOT a port of the sm80 mma.sync kernels: the problem is computed transposed,
//     Y^T[N_out, tokens] = W^T[N_out, K] . X^T[K, tokens]
// so the streamed operand (weights) sits on the UMMA M side (128 output features per CTA) and the tiny
// activation tile is UMMA N (16..128 tokens).  Per 128-wide K group (= one AWQ quant group):
//   * a producer thread issues two 1-D TMA bulk copies: one contiguous 8.5 KB block of packed u4 weights +
//     their fused (scale, -zero*scale) pairs, and the contiguous activation tile (already in UMMA
//     K-major "interleaved" layout in global memory);
//   * 8 dequant warps unpack u4 -> f16 with lop3 and apply w = fma(q, s, t) in f16 (bit-identical to the
//     reference's transform, transform.h:48-74 / cast.cu:155) and write the A operand straight into
//     TENSOR MEMORY with tcgen05.st (no shared-memory round trip for the dequantised weights);
//   * one thread issues tcgen05.mma (A from TMEM, B from smem), fp32 accumulators in TMEM;
//   * the epilogue reads accumulators with tcgen05.ld, does the deterministic split-K fix-up through an
//     fp32 workspace (last-arriving CTA reduces in split order), bias / gated-SiLU, and writes fp16/bf16.
// The kernel is HBM-bound at decode batch sizes: algorithmic bytes = packed weight stream, read once.
#pragma once

#include "umma.cuh"

namespace b200 {
namespace gemm {

constexpr int kBN        = 128;  // output features per CTA (UMMA M)
constexpr int kKG        = 128;  // K per pipeline stage (= quant group size)
constexpr int kMaxStages = 4;
constexpr int kThreads   = 320;  // warp0 TMA, warp1 MMA/TMEM, warps2-5 dequant A + epilogue, warps6-9 dequant B
// TMEM budget: D accumulators occupy columns [0, MT); the dequantised A operand uses 64 columns per stage starting
// at column a_col.  MT <= 64 with 3 stages fits 256 columns, so two CTAs can be resident per SM.

enum WeightFormat : int
{
    kDense = 0,  // T weights, [k/8][128][8] tiles
    kU4    = 1,  // AWQ-style u4 + (scale, zero) per 128-group
    kFp8   = 2,  // e4m3 + per-128x128-block scale
};

constexpr int kU4BlockBytes    = 8192 + 512;           // one (n-tile, k-group): data + fused params
constexpr int kDenseBlockBytes = kBN * kKG * 2;        // 32 KB
constexpr int kFp8BlockBytes   = kBN * kKG + 16;       // data + block scale (T, padded to 16 B)

// Development aid (tools/gemm_trace.cu compiles this TU with -DB200_GEMM_TRACE): per-role clock64 stamps of the
// first kTraceCtas CTAs.  Absent from the product build.
#ifdef B200_GEMM_TRACE
constexpr int kTraceCtas = 8, kTraceRoles = 6, kTraceSlots = 128;
__device__ long long g_trace[kTraceCtas][kTraceRoles][kTraceSlots];
#define GEMM_TRACE(role, slot)                                                                                         \
    do {                                                                                                               \
        const int _c = blockIdx.y * gridDim.x + blockIdx.x;                                                            \
        if (_c < kTraceCtas && (slot) < kTraceSlots) g_trace[_c][role][slot] = clock64();                              \
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
#define GEMM_WAIT(bar, parity, tag, it) mbar_wait(bar, parity)
#endif

struct Params {
    const uint8_t* w;   // prepared weight stream
    const void*    xp;  // packed activations [K/8][MT][8] T
    void*          y;
    const void*    bias;      // [N] T or null
    int            N, K;
    int            MT;  // padded token count (multiple of 16, <= 128)
    int            M;   // real tokens
    int            splits;  // K splits = thread-block cluster size along grid.y (1, 2, 4 or 8)
    int            stages;
    int            epilogue;  // 0 none, 1 gated SiLU (columns interleaved gate/up -> N/2 outputs)
    int            out_mode;  // 0 row-major [M][ldy]; 1 packed [Nout/8][out_mt][8]
    int            ldy;
    int            out_mt;
    int            tmem_cols;  // 256 or 512 (power of two)
    int            a_col;      // first TMEM column of the A-operand ring
};

B200_DEVICE int w_block_bytes(int fmt)
{
    return fmt == kU4 ? kU4BlockBytes : (fmt == kDense ? kDenseBlockBytes : kFp8BlockBytes);
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
            const float silu = acc / (1.f + expf(-acc));
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
        const float silu = a0 / (1.f + expf(-a0));
        store_out<T>(p, tok, n_g >> 1, silu * a1);
    }
    else {
        T* y = reinterpret_cast<T*>(p.y);
        T* d = p.out_mode == 0 ? y + (size_t)tok * p.ldy + n_g : y + ((size_t)(n_g >> 3) * p.out_mt + tok) * 8 + (n_g & 7);
        *reinterpret_cast<uint32_t*>(d) = Ops::pack2(a0, a1);
    }
}

B200_DEVICE uint32_t mapa_shared(uint32_t addr, int cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
B200_DEVICE void st_shared_cluster_f32(uint32_t addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
B200_DEVICE void cluster_arrive_release()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
B200_DEVICE void cluster_wait_acquire()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template<class T, int FMT>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const Params p)
{
    using Ops = TypeOps<T>;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t                       bar_full[kMaxStages], bar_empty[kMaxStages], bar_aready[kMaxStages];
    __shared__ uint64_t                       bar_dready;
    __shared__ uint32_t                       s_tmem;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform
    const int nt = blockIdx.x, sp = blockIdx.y;
    const int kgs = p.K / kKG;
    const int kg0 = kgs * sp / p.splits, kg1 = kgs * (sp + 1) / p.splits;
    const int nk = kg1 - kg0;
    if (threadIdx.x == 64) GEMM_TRACE(4, 0);

    constexpr int WB     = FMT == kU4 ? kU4BlockBytes : (FMT == kDense ? kDenseBlockBytes : kFp8BlockBytes);
    const int     XB     = p.MT * 256;  // 16 k8-chunks x MT x 16 B
    const int     WB_pad = (WB + 1023) & ~1023;
    const int     stage_bytes = WB_pad + XB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
            mbar_init(smem_u32(&bar_aready[s]), 256);
        }
        mbar_init(smem_u32(&bar_dready), 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        umma::tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (p.splits > 1) {
        // every CTA of the cluster must be running before anyone stores into its shared memory
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    if (threadIdx.x == 64) GEMM_TRACE(4, 1);

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            const uint64_t pol_w = make_evict_first_policy();
            int      s  = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nk; ++i) {
                GEMM_WAIT(smem_u32(&bar_empty[s]), ph ^ 1, 1, i);
                GEMM_TRACE(0, i);
                const uint32_t bar = smem_u32(&bar_full[s]);
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_expect_tx(bar, WB + XB);
                bulk_g2s_hint(dst, p.w + ((size_t)nt * kgs + kg0 + i) * WB, WB, bar, pol_w);
                bulk_g2s(dst + WB_pad, reinterpret_cast<const uint8_t*>(p.xp) + (size_t)(kg0 + i) * XB, XB, bar);
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    }
    else if (warp == 1) {
        // ================================ MMA issuer ================================
        // The whole warp runs this loop under warp-uniform control flow and one elected lane issues; every operand
        // is derived from uniform values, so ptxas keeps descriptors in uniform registers and a k-step costs two
        // uniform adds + UTCHMMA (a lane-0 branch makes it rebuild and R2UR every descriptor: ~100 cycles per MMA,
        // which made the issue thread, not HBM, the bound of the whole kernel).
        const uint32_t idesc  = umma::make_idesc_f16(kBN, p.MT, std::is_same<T, nv_bfloat16>::value);
        const uint32_t b_step = (uint32_t)(2 * p.MT * 16) >> 4;  // descriptor start-address units per k-step
        const uint32_t a_step = (uint32_t)(2 * kBN * 16) >> 4;
        const uint32_t smem0  = smem_u32(smem);
        int            s      = 0;
        uint32_t       ph     = 0;
        for (int i = 0; i < nk; ++i) {
            GEMM_WAIT(smem_u32(&bar_full[s]), ph, 2, i);
            if (FMT != kDense) {
                GEMM_WAIT(smem_u32(&bar_aready[s]), ph, 3, i);
            }
            umma::fence_after_sync();
            if (lane == 0) GEMM_TRACE(3, i);
            const uint32_t sbase = smem0 + (uint32_t)s * (uint32_t)stage_bytes;
            const uint64_t bd0   = umma::make_smem_desc(sbase + WB_pad, p.MT * 16, 128);
            const uint64_t ad0   = umma::make_smem_desc(sbase, kBN * 16, 128);
            const uint32_t a_tm  = tmem + p.a_col + s * 64;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kKG / 16; ++ks) {
                    const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                    if (FMT == kDense) {
                        umma::mma_f16_ss(tmem, ad0 + (uint64_t)(ks * a_step), bd0 + (uint64_t)(ks * b_step), idesc, acc);
                    }
                    else {
                        umma::mma_f16_ts(tmem, a_tm + ks * 8, bd0 + (uint64_t)(ks * b_step), idesc, acc);
                    }
                }
                umma::commit(smem_u32(&bar_empty[s]));
            }
            __syncwarp();
            if (lane == 0) GEMM_TRACE(5, i);
            if (++s == p.stages) {
                s = 0;
                ph ^= 1;
            }
        }
        if (elect_one()) {
            umma::commit(smem_u32(&bar_dready));
        }
        __syncwarp();
    }
    else {
        // ================================ dequant warps ================================
        const int      grp       = (warp - 2) >> 2;              // 0: warps 2-5, 1: warps 6-9
        const int      lane_base = 32 * (warp & 3);              // TMEM lanes this warp may touch
        const int      n_local   = lane_base + lane;
        const uint32_t trow      = tmem + ((uint32_t)lane_base << 16);
        if constexpr (FMT == kU4) {
            // Every dequant warp takes part in EVERY stage (warps 2-5: k-chunk 0, warps 6-9: k-chunk 1 of the
            // 128-wide group), so each thread observes every phase of bar_full[s] (a parity wait is ambiguous for
            // a waiter that skips phases) and the per-stage dequant latency is halved.
            const int c  = grp;
            int       s  = 0;
            uint32_t  ph = 0;
            for (int i = 0; i < nk; ++i) {
                GEMM_WAIT(smem_u32(&bar_full[s]), ph, 4, i);
                if ((warp == 2 || warp == 6) && lane == 0) GEMM_TRACE(1, i);
                const uint32_t wsm = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t st2 = lds32(wsm + 8192 + n_local * 4);
                const uint32_t s2 = __byte_perm(st2, st2, 0x1010), t2 = __byte_perm(st2, st2, 0x3232);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint4 v = lds128(wsm + ((c * 2 + hh) * 128 + n_local) * 16);
                    uint32_t    r[16];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t h[4];
                        Ops::cvt_u4x8((&v.x)[u], h);
#pragma unroll
                        for (int x = 0; x < 4; ++x) r[u * 4 + x] = Ops::fma2(h[x], s2, t2);
                    }
                    umma::st_32x32b_x16(trow + p.a_col + s * 64 + c * 32 + hh * 16, r);
                }
                umma::wait_st();
                umma::fence_before_sync();
                mbar_arrive(smem_u32(&bar_aready[s]));
                if ((warp == 2 || warp == 6) && lane == 0) GEMM_TRACE(2, i);
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        // ================================ epilogue, part 1 (all 8 warps) ================================
        // warps 2-5 take tokens [0, MT/2), warps 6-9 tokens [MT/2, MT) of the 32 accumulator rows their TMEM lane
        // quarter holds.  splits == 1: straight to the output.  splits > 1 (thread-block cluster along K): each
        // value goes to the shared memory of the cluster CTA that owns its token (reduce-scatter over DSMEM).
        GEMM_WAIT(smem_u32(&bar_dready), 0, 5, 0);
        umma::fence_after_sync();
        if (threadIdx.x == 64) GEMM_TRACE(4, 2);
        const int n_g  = nt * kBN + n_local;
        const int half = p.MT >> 1;
        if (p.splits == 1) {
            for (int t0 = grp * half; t0 < (grp + 1) * half; t0 += 8) {
                uint32_t v[8];
                umma::ld_32x32b_x8(trow + t0, v);
                umma::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) finalize<T>(p, t0 + i, n_g, __uint_as_float(v[i]));
            }
        }
        else {
            const int      tps = p.MT / p.splits;  // tokens owned per cluster CTA
            const uint32_t red = smem_u32(smem) + (uint32_t)p.stages * (uint32_t)stage_bytes;
            // my slot in every owner's buffer: red[src = sp][tl][n]
            const uint32_t mine = red + (uint32_t)((sp * tps) * kBN + n_local) * 4u;
            for (int t0 = grp * half; t0 < (grp + 1) * half; t0 += 8) {
                uint32_t v[8];
                umma::ld_32x32b_x8(trow + t0, v);
                umma::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int      t     = t0 + i;
                    const int      owner = t / tps;
                    const uint32_t dst   = mapa_shared(mine + (uint32_t)((t - owner * tps) * kBN) * 4u, owner);
                    st_shared_cluster_f32(dst, __uint_as_float(v[i]));
                }
            }
        }
    }
    if (threadIdx.x == 64) GEMM_TRACE(4, 3);
    if (p.splits > 1) {
        // ================================ epilogue, part 2: owner-side ordered sum ================================
        cluster_arrive_release();
        cluster_wait_acquire();
        if (warp >= 2) {
            const int      tps  = p.MT / p.splits;
            const int      tid  = threadIdx.x - 64;  // 0..255
            const uint8_t* redp = smem + (size_t)p.stages * stage_bytes;
            for (int idx = tid; idx < tps * (kBN / 2); idx += 256) {
                const int tl = idx / (kBN / 2), np = idx % (kBN / 2);
                float2    acc = make_float2(0.f, 0.f);
                for (int r = 0; r < p.splits; ++r) {  // fixed order: deterministic
                    const float2 v = *reinterpret_cast<const float2*>(redp + ((size_t)(r * tps + tl) * kBN + 2 * np) * 4);
                    acc.x += v.x;
                    acc.y += v.y;
                }
                finalize_pair<T>(p, sp * tps + tl, nt * kBN + 2 * np, acc.x, acc.y);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        umma::tmem_dealloc(tmem, p.tmem_cols);
    }
}

}  // namespace gemm
}  // namespace b200
