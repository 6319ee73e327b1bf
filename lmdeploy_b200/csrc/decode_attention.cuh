// Paged-KV flash-decoding attention for sm_100a (B200).
//
// Replaces, behind the C ABI in include/b200_turbomind.h, the reference's
//   dispatchDecoding<T>(AttentionParams<T>)            src/turbomind/kernels/attention/decoding.cu:13-39
//   attention_kernel<Decoding_MMA<...>>                attention_universal.h:355-553, impl_81616.h
//   attention::reduce / invokeReduceV3                 reduce.cu:13-298
// It is NOT a port: the design is a persistent, warp-specialised kernel
//   * one producer warp streams KV page tiles (64 tokens of one kv-head) from the block table into a
//     shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * four consumer warps de-quantise int8/int4 in registers straight into mma.sync fragments
//     (K as A-operand of S = K.Q^T, V^T as A-operand of O^T = V^T.P), online base-2 softmax;
//   * work is a flat, balanced partition of all (sequence, kv-head, tile) units over a fixed grid
//     ("stream-K"); a segment that straddles CTAs leaves fp32 partials that a fix-up kernel merges in
//     a fixed order (deterministic run to run).
// The kernel is HBM-bound (every KV byte is read exactly once); tensor cores are used only because
// CUDA-core FMA throughput cannot keep up with 6.5 TB/s of int8/int4 operands.
//
// Shared-memory tile layout: rows (tokens) are copied in groups of 16 with 32 bytes of padding between
// groups; together with the slot->token permutation below this makes every LDS of the fragment loads
// bank-conflict free for 64/128/256-byte rows without TMA swizzling (pages are addressed by pointer,
// so no tensor map exists for them).
#pragma once

#include "common.cuh"

namespace b200 {
namespace attn {

constexpr int kHeadDim       = 128;
constexpr int kTile          = 64;  // tokens per pipeline stage
constexpr int kGroupRows     = 16;
constexpr int kGroupPad      = 32;
constexpr int kConsumerWarps = 4;
constexpr int kThreads       = (kConsumerWarps + 1) * 32;
constexpr int kMaxGroup      = 8;  // q heads per kv head handled by one CTA (mma N)

template<int KVBITS>
struct TileCfg {
    static constexpr bool kQuant     = KVBITS != 16;
    static constexpr int  kRowBytes  = kHeadDim * KVBITS / 8;
    static constexpr int  kGroupSize = kGroupRows * kRowBytes + kGroupPad;
    static constexpr int  kTileBytes = (kTile / kGroupRows) * kGroupSize;
    static constexpr int  kParamBytes = kQuant ? kTile * 4 : 0;  // (scale, zero) as T per token
    static constexpr int  kStageBytes = 2 * kTileBytes + 2 * kParamBytes;
    // offsets inside a stage
    static constexpr int kOffK  = 0;
    static constexpr int kOffV  = kTileBytes;
    static constexpr int kOffKP = 2 * kTileBytes;
    static constexpr int kOffVP = 2 * kTileBytes + kParamBytes;
};

// ------------------------------------------------------------------------------------------------
// RoPE constants: init_rope_kernel_param (src/turbomind/models/attention_weight.cc:37-93)
// ------------------------------------------------------------------------------------------------
struct RopeConst {
    int   type;
    int   dim;
    float scale_factor;
    float inv_factor;
    float attention_scaling;
    float llama3_alpha, llama3_beta;
    float yarn_div2, yarn_mulmin;
    int   exact;  // b200_rope_config::exact_math
};


__device__ __forceinline__ float mul_ftz(float a, float b)
{
    float c;
    asm("mul.ftz.f32 %0, %1, %2;" : "=f"(c) : "f"(a), "f"(b));
    return c;
}
__device__ __forceinline__ float fma_ftz(float a, float b, float c)
{
    float d;
    asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// inv_freq of the pair starting at even dim `idx` (rotary_embedding.h:10-52).
// Default: the instruction sequence of the reference's release build (--use_fast_math: ex2.approx.ftz, ftz
// multiplies, mul+add contracted to fma) so results are bit-identical to TurboMind on the same GPU; `exact`:
// correctly rounded exp2 and un-fused fp32 ops (what the CPU oracle computes).
__device__ __forceinline__ float rope_inv_freq(const RopeConst& r, int idx)
{
    if (r.exact) {
        const float e    = __fmul_rn((float)idx, r.scale_factor);
        const float freq = (float)exp2((double)e);
        if (r.type == 5) {
            float sm = fmaf(r.llama3_alpha, freq, -r.llama3_beta);
            sm       = fmaxf(0.f, fminf(1.f, sm));
            const float a = __fmul_rn(__fsub_rn(1.f, sm), freq);
            const float b = __fmul_rn(a, r.inv_factor);
            const float c = __fmul_rn(sm, freq);
            return __fadd_rn(b, c);
        }
        if (r.type == 4) {
            float alpha = __fsub_rn(__fmul_rn((float)idx, r.yarn_div2), r.yarn_mulmin);
            alpha       = fmaxf(0.f, fminf(1.f, alpha));
            const float t = __fmul_rn(__fmul_rn(freq, alpha), __fsub_rn(1.f, r.inv_factor));
            return __fsub_rn(freq, t);
        }
        return __fmul_rn(r.inv_factor, freq);
    }
    const float freq = exp2_approx(mul_ftz((float)idx, r.scale_factor));
    if (r.type == 5) {
        // (1 - smooth) * freq * inv_factor + smooth * freq
        float sm = fma_ftz(r.llama3_alpha, freq, -r.llama3_beta);
        sm       = fmaxf(0.f, fminf(1.f, sm));
        const float a = mul_ftz(mul_ftz(1.f - sm, freq), r.inv_factor);
        return fma_ftz(sm, freq, a);
    }
    if (r.type == 4) {
        // freq - freq * alpha * (1 - inv_factor)
        float alpha = fma_ftz((float)idx, r.yarn_div2, -r.yarn_mulmin);
        alpha       = fmaxf(0.f, fminf(1.f, alpha));
        const float t = mul_ftz(freq, alpha);
        return fma_ftz(-t, 1.f - r.inv_factor, freq);
    }
    return mul_ftz(r.inv_factor, freq);
}

__device__ __forceinline__ float2 rope_cos_sin(const RopeConst& r, int idx, int pos)
{
    if (r.exact) {
        const float ang = __fmul_rn((float)pos, rope_inv_freq(r, idx));
        double      s, c;
        sincos((double)ang, &s, &c);
        return make_float2(__fmul_rn((float)c, r.attention_scaling), __fmul_rn((float)s, r.attention_scaling));
    }
    // sincosf under --use_fast_math = sin.approx.ftz / cos.approx.ftz of the ftz product (no range reduction)
    const float ang = mul_ftz((float)pos, rope_inv_freq(r, idx));
    float       s, c;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
    return make_float2(mul_ftz(c, r.attention_scaling), mul_ftz(s, r.attention_scaling));
}


// bias + RoPE of 4 consecutive dims [di, di + 4) of one head vector, T arithmetic exactly as the reference binary's
// (rotary_embedding.h:168-181 compiled with --use_fast_math: `c*x0 - s*x1` contracted to HMUL2 + HFMA2)
template<class T>
__device__ __forceinline__ void rope_rotate4(T (&x)[4], const float2 (&cs)[2], const RopeConst& rope, int di)
{
    using Ops = TypeOps<T>;
    if (rope.type == 0) {
        return;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (di + 2 * i < rope.dim) {
            const T c = Ops::from_f32(cs[i].x), s = Ops::from_f32(cs[i].y);
            const T a = x[2 * i], bb = x[2 * i + 1];
            x[2 * i]     = Ops::fma(c, a, Ops::neg(Ops::mul(s, bb)));
            x[2 * i + 1] = Ops::fma(c, bb, Ops::mul(s, a));
        }
    }
}

// out of line: the exact-math variant drags double-precision sincos (and its stack frame) into whoever inlines it
__device__ __noinline__ float2 rope_cos_sin_call(const RopeConst& r, int idx, int pos)
{
    return rope_cos_sin(r, idx, pos);
}

// (cos, sin) of the two pairs starting at even dim `di` for position `ti`: from the per-step table when there is one
__device__ __forceinline__ void rope_cs2(float2 (&cs)[2], const RopeConst& rope, const float2* table, int tok, int di, int ti)
{
    if (table) {
        const float4 v = *reinterpret_cast<const float4*>(table + (size_t)tok * (kHeadDim / 2) + (di >> 1));
        cs[0]          = make_float2(v.x, v.y);
        cs[1]          = make_float2(v.z, v.w);
    }
    else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            cs[i] = (rope.type != 0 && di + 2 * i < rope.dim) ? rope_cos_sin_call(rope, di + 2 * i, ti) : make_float2(1.f, 0.f);
        }
    }
}

template<class T>
__device__ __forceinline__ void load4(const T* p, T (&x)[4])
{
    uint2 u                          = *reinterpret_cast<const uint2*>(p);
    *reinterpret_cast<uint2*>(&x[0]) = u;
}

// Quantise one K or V head vector spread over a warp (lane = 4 consecutive dims): warp_stats + ConvertKvCache
// (quantization.h:341-366, 427-490).  Returns this lane's packed bytes in `bytes` (int8: 4 bytes; int4: the even lane of a pair
// holds 4 bytes covering 8 dims, valid = false on odd lanes) and the (scale, zero) pair as two T in one u32.
template<class T, int KVBITS>
__device__ __forceinline__ void quantize_head(const T (&x)[4], int lane, uint32_t& bytes, bool& valid, uint32_t& param)
{
    using Ops = TypeOps<T>;
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
    if constexpr (KVBITS == 8) {
        bytes = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
        valid = true;
    }
    else {
        // nibble j of each u32 holds element [0,2,4,6,1,3,5,7][j]  (ConvertKvCache<T,uint4_t>::pack)
        const uint32_t mine = (lane & 1) ? ((q[0] << 8) | (q[2] << 12) | (q[1] << 24) | (q[3] << 28)) :
                                           ((q[0]) | (q[2] << 4) | (q[1] << 16) | (q[3] << 20));
        const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
        bytes = mine | other;
        valid = (lane & 1) == 0;
    }
    T pr[2] = {scale, zero};
    param   = *reinterpret_cast<uint32_t*>(&pr[0]);
}

struct KernelParams {
    const void* q_rot;  // [token][Hq][D]  (bias + RoPE applied)
    void*       out;    // [token][Hq*D], row stride out_stride elements
    int64_t     out_stride;
    char**      block_ptrs;
    const int*  cu_block_nums;  // [B+1]
    const int*  cu_k_len;       // [B+1]
    const int*  tile_prefix;    // [B+1] prefix sum of tiles per sequence (plan kernel)
    float*      partial_O;      // [2*grid][8][128]
    float*      partial_ML;     // [2*grid][8][2]
    int         layer_offset;   // byte offset of this layer inside a block
    int         block_len;
    int         batch;
    int         num_heads;
    int         num_kv_heads;
    float       qk_scale;  // log2(e)/sqrt(D)  (AttentionParams::inv_sqrt_dh)
    int         out_mt;    // > 0: write `out` in the UMMA K-major tile layout [Hq*D/8][out_mt][8]
    // ---- in-kernel prologue (fused & 1): bias + RoPE of q at segment start, and the warp that owns the new token's slot
    // quantises K/V, stores them into the page and patches its shared-memory tile (attention_universal.h:168-348 does the
    // same inside the reference kernel); q_rot is then unused
    int            fused;  // bit 0: in-kernel prologue, bit 1: in-kernel fix-up (the CTA that completes a straddling
                           // (sequence, kv-head) last merges its partials; no reduce kernel)
    const void*    q;
    const void*    k;
    const void*    v;
    int64_t        stride;
    const void*    q_bias;
    const void*    k_bias;
    const void*    v_bias;
    const int*     cu_q_len;
    const int*     readonly_block_num;
    const float2*  rope_table;
    RopeConst      rope;
    int*           seg_counters;  // [batch * Hkv] arrivals per straddling segment; zero on entry, left zero
    int            grid_main;
    int            early;  // the producer warp streams KV tiles before griddepcontrol.wait (see the kernel's head)
};

// slot (0..15 inside a warp's 16-key MMA) -> token index inside the 64-token tile
B200_DEVICE int slot_token(int c, int warp)
{
    // group table G = [0,2,1,3,2,0,3,1] packed 2 bits each
    const int grp = (0x72D8 >> (2 * (c & 7))) & 3;
    const int i   = (c >> 3) + 2 * ((c & 7) >> 2);
    return 16 * grp + 4 * warp + i;
}

template<int KVBITS>
B200_DEVICE int token_smem_offset(int tok)
{
    return (tok >> 4) * TileCfg<KVBITS>::kGroupSize + (tok & 15) * TileCfg<KVBITS>::kRowBytes;
}

// first dim of the 4 consecutive dims lane-class t contracts in k-step j of S = K.Q^T
template<int KVBITS>
B200_DEVICE int qk_dim0(int t, int j)
{
    if constexpr (KVBITS == 8) {
        return 16 * t + 64 * (j >> 2) + 4 * (j & 3);
    }
    else if constexpr (KVBITS == 4) {
        return 32 * t + 4 * j;
    }
    else {
        return 8 * (t + 4 * (j >> 1)) + 4 * (j & 1);
    }
}

// output dim held by lane-group g in m-tile j, element e (0: m-slot g, 1: m-slot g+8)
template<int KVBITS>
B200_DEVICE int pv_dim(int g, int j, int e)
{
    if constexpr (KVBITS == 16) {
        return (j < 4 ? 8 * g + 2 * j : 64 + 8 * g + 2 * (j - 4)) + e;
    }
    else {
        return 16 * g + 2 * j + e;
    }
}

struct WorkIter {
    int        batch, hkv;
    const int* prefix;
    int        b, h, tile, tiles;
    B200_DEVICE void seek(long long idx)
    {
        // find b: hkv*prefix[b] <= idx < hkv*prefix[b+1]
        int lo = 0, hi = batch;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if ((long long)hkv * prefix[mid] <= idx) {
                lo = mid;
            }
            else {
                hi = mid;
            }
        }
        b     = lo;
        tiles = prefix[b + 1] - prefix[b];
        // lower_bound above may land on an empty sequence only if idx is past the end
        long long r = idx - (long long)hkv * prefix[b];
        h           = tiles ? (int)(r / tiles) : 0;
        tile        = tiles ? (int)(r % tiles) : 0;
    }
    B200_DEVICE void next()
    {
        if (++tile == tiles) {
            tile = 0;
            if (++h == hkv) {
                h = 0;
                do {
                    ++b;
                    tiles = (b < batch) ? prefix[b + 1] - prefix[b] : 1;
                } while (tiles == 0);
            }
        }
    }
};

// The main-kernel CTAs that hold a piece of segment (sequence b, kv-head h).  CTA c owns units [total * c / grid,
// total * (c + 1) / grid).  With total >= grid every CTA is non-empty and the pieces sit in CTAs cf..cl; with total < grid a
// CTA holds at most ONE unit (= one tile), empty CTAs lie in between, and the pieces are the segment's `tiles` units, unit
// u living in CTA ((u + 1) * grid - 1) / total.  (Round 1 walked cf..cl in both cases and read the never-written -- or,
// worse, stale -- partial slots of the empty CTAs in between.)
struct SegCtas {
    int       cf, cl, first_slot, n;  // n = number of pieces
    int       sparse;
    long long s0, total;
};

template<class T, int KVBITS, int STAGES>
struct SharedStorage {
    using Cfg = TileCfg<KVBITS>;
    alignas(128) uint8_t stage[STAGES][Cfg::kStageBytes];
    alignas(16) float merge_O[2][kMaxGroup][kHeadDim];
    float    merge_M[kConsumerWarps][kMaxGroup];
    float    merge_L[kConsumerWarps][kMaxGroup];
    float    final_M[kMaxGroup];
    float    final_L[kMaxGroup];
    SegCtas  seg;       // in-kernel fix-up: where the pieces of the segment being closed live
    int      fix_last;  // ... and whether this CTA arrived last (it merges)
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
};

// swizzled index into the [head][128] fp32 merge rows: keeps (d, d+1) pairs adjacent and spreads the 16 lanes of
// a half-warp over 16 distinct 8-byte slots (bank-conflict-free 64-bit accesses in the segment merge)
B200_DEVICE int merge_idx(int hd, int d)
{
    return hd * kHeadDim + (d ^ ((((d >> 5) & 3) << 1) | ((hd >> 1) << 3)));
}

template<class T, int KVBITS>
struct ConsumerCtx {
    int      warp, g, t;
    int      tokA, tokB, tokV0, tokV1, tokV2, tokV3;
    uint32_t offA, offB, offV0, offV1, offV2, offV3;
};

// one 64-token tile: this warp's 16 keys.  `st` = shared address of the stage.
template<class T, int KVBITS>
B200_DEVICE void process_tile(const ConsumerCtx<T, KVBITS>& cx, uint32_t st, bool residue, int k_len, int t0, float c,
                              const uint32_t (&qf)[8][2], float (&O)[8][4], float (&M)[2], float (&L)[2])
{
    using Ops = TypeOps<T>;
    using Cfg = TileCfg<KVBITS>;
    const int g = cx.g, t = cx.t;
    const int tokA = cx.tokA, tokB = cx.tokB, tokV0 = cx.tokV0, tokV1 = cx.tokV1, tokV2 = cx.tokV2, tokV3 = cx.tokV3;
    const uint32_t offA = cx.offA, offB = cx.offB, offV0 = cx.offV0, offV1 = cx.offV1, offV2 = cx.offV2, offV3 = cx.offV3;
    // ------------------------------ S = K . Q^T ------------------------------
    float S[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (KVBITS == 8) {
        const uint32_t pA  = lds32(st + Cfg::kOffKP + tokA * 4);
        const uint32_t pB  = lds32(st + Cfg::kOffKP + tokB * 4);
        const uint32_t sA2 = __byte_perm(pA, pA, 0x1010), zA2 = __byte_perm(pA, pA, 0x3232);
        const uint32_t sB2 = __byte_perm(pB, pB, 0x1010), zB2 = __byte_perm(pB, pB, 0x3232);
        uint4          ka[2], kb[2];
        ka[0] = lds128(st + Cfg::kOffK + offA + 16 * t);
        ka[1] = lds128(st + Cfg::kOffK + offA + 16 * (t + 4));
        kb[0] = lds128(st + Cfg::kOffK + offB + 16 * t);
        kb[1] = lds128(st + Cfg::kOffK + offB + 16 * (t + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t ua = (&ka[j >> 2].x)[j & 3];
            const uint32_t ub = (&kb[j >> 2].x)[j & 3];
            uint32_t       a[4];
            Ops::cvt_u8x4(ua, a[0], a[2]);
            Ops::cvt_u8x4(ub, a[1], a[3]);
            a[0] = Ops::fma2(a[0], sA2, zA2);
            a[2] = Ops::fma2(a[2], sA2, zA2);
            a[1] = Ops::fma2(a[1], sB2, zB2);
            a[3] = Ops::fma2(a[3], sB2, zB2);
            Ops::mma16816(S, a, qf[j]);
        }
    }
    else if constexpr (KVBITS == 4) {
        const uint32_t pA  = lds32(st + Cfg::kOffKP + tokA * 4);
        const uint32_t pB  = lds32(st + Cfg::kOffKP + tokB * 4);
        const uint32_t sA2 = __byte_perm(pA, pA, 0x1010), zA2 = __byte_perm(pA, pA, 0x3232);
        const uint32_t sB2 = __byte_perm(pB, pB, 0x1010), zB2 = __byte_perm(pB, pB, 0x3232);
        const uint4    ka  = lds128(st + Cfg::kOffK + offA + 16 * t);
        const uint4    kb  = lds128(st + Cfg::kOffK + offB + 16 * t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t ha[4], hb[4];
            Ops::cvt_u4x8((&ka.x)[i], ha);
            Ops::cvt_u4x8((&kb.x)[i], hb);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                ha[x] = Ops::fma2(ha[x], sA2, zA2);
                hb[x] = Ops::fma2(hb[x], sB2, zB2);
            }
            // u32 i covers dims 8i..8i+7 of this lane's chunk: k-steps 2i (dims 0-3) and 2i+1 (dims 4-7)
            uint32_t a0[4] = {ha[0], hb[0], ha[1], hb[1]};
            uint32_t a1[4] = {ha[2], hb[2], ha[3], hb[3]};
            Ops::mma16816(S, a0, qf[2 * i]);
            Ops::mma16816(S, a1, qf[2 * i + 1]);
        }
    }
    else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 ka = lds128(st + Cfg::kOffK + offA + 16 * (t + 4 * i));
            const uint4 kb = lds128(st + Cfg::kOffK + offB + 16 * (t + 4 * i));
            uint32_t    a0[4] = {ka.x, kb.x, ka.y, kb.y};
            uint32_t    a1[4] = {ka.z, kb.z, ka.w, kb.w};
            Ops::mma16816(S, a0, qf[2 * i]);
            Ops::mma16816(S, a1, qf[2 * i + 1]);
        }
    }

    // ------------------------------ online softmax ------------------------------
    // S[0],S[1]: row g (tokA) heads 2t,2t+1;  S[2],S[3]: row g+8 (tokB)
    const bool maskA = residue && (t0 + tokA >= k_len);
    const bool maskB = residue && (t0 + tokB >= k_len);
    if (maskA) {
        S[0] = S[1] = -INFINITY;
    }
    if (maskB) {
        S[2] = S[3] = -INFINITY;
    }
    float mx[2] = {fmaxf(S[0], S[2]), fmaxf(S[1], S[3])};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 4));
        mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 8));
        mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 16));
    }
    float newM[2] = {fmaxf(M[0], mx[0]), fmaxf(M[1], mx[1])};
    // a whole 16-key slice can only be fully masked in a residue tile
    const bool dead0 = newM[0] == -INFINITY, dead1 = newM[1] == -INFINITY;
    if (newM[0] != M[0] || newM[1] != M[1]) {
        const float r0 = dead0 ? 1.f : exp2_approx((M[0] - newM[0]) * c);
        const float r1 = dead1 ? 1.f : exp2_approx((M[1] - newM[1]) * c);
        L[0] *= r0;
        L[1] *= r1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            O[j][0] *= r0;
            O[j][1] *= r1;
            O[j][2] *= r0;
            O[j][3] *= r1;
        }
        M[0] = newM[0];
        M[1] = newM[1];
    }
    const float mc0 = dead0 ? 0.f : M[0] * c, mc1 = dead1 ? 0.f : M[1] * c;
    float       P[4];
    P[0] = exp2_approx(fmaf(S[0], c, -mc0));
    P[1] = exp2_approx(fmaf(S[1], c, -mc1));
    P[2] = exp2_approx(fmaf(S[2], c, -mc0));
    P[3] = exp2_approx(fmaf(S[3], c, -mc1));
    L[0] += P[0] + P[2];  // per-lane partial row sums; reduced across g at segment end
    L[1] += P[1] + P[3];

    uint32_t pf[2];
    pf[0] = transpose_8x8_b16(Ops::pack2(P[0], P[1]));  // k-slots 2t,2t+1 ; head g
    pf[1] = transpose_8x8_b16(Ops::pack2(P[2], P[3]));  // k-slots 2t+8,2t+9 ; head g

    // ------------------------------ O^T += V^T . P ------------------------------
    if constexpr (KVBITS == 8) {
        uint32_t p0 = lds32(st + Cfg::kOffVP + tokV0 * 4), p1 = lds32(st + Cfg::kOffVP + tokV1 * 4);
        uint32_t p2 = lds32(st + Cfg::kOffVP + tokV2 * 4), p3 = lds32(st + Cfg::kOffVP + tokV3 * 4);
        if (residue) {
            if (t0 + tokV0 >= k_len) p0 = 0;
            if (t0 + tokV1 >= k_len) p1 = 0;
            if (t0 + tokV2 >= k_len) p2 = 0;
            if (t0 + tokV3 >= k_len) p3 = 0;
        }
        const uint32_t s01 = __byte_perm(p0, p1, 0x5410), z01 = __byte_perm(p0, p1, 0x7632);
        const uint32_t s23 = __byte_perm(p2, p3, 0x5410), z23 = __byte_perm(p2, p3, 0x7632);
        const uint4    v0 = lds128(st + Cfg::kOffV + offV0 + 16 * g), v1 = lds128(st + Cfg::kOffV + offV1 + 16 * g);
        const uint4    v2 = lds128(st + Cfg::kOffV + offV2 + 16 * g), v3 = lds128(st + Cfg::kOffV + offV3 + 16 * g);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const uint32_t A = (&v0.x)[jj], Bq = (&v1.x)[jj], C = (&v2.x)[jj], Dq = (&v3.x)[jj];
            const uint32_t X = __byte_perm(A, Bq, 0x5140), Y = __byte_perm(A, Bq, 0x7362);
            const uint32_t X2 = __byte_perm(C, Dq, 0x5140), Y2 = __byte_perm(C, Dq, 0x7362);
            uint32_t       a[4];
            Ops::cvt_u8x4(X, a[0], a[1]);
            Ops::cvt_u8x4(X2, a[2], a[3]);
            a[0] = Ops::fma2(a[0], s01, z01);
            a[1] = Ops::fma2(a[1], s01, z01);
            a[2] = Ops::fma2(a[2], s23, z23);
            a[3] = Ops::fma2(a[3], s23, z23);
            Ops::mma16816(O[2 * jj], a, pf);
            Ops::cvt_u8x4(Y, a[0], a[1]);
            Ops::cvt_u8x4(Y2, a[2], a[3]);
            a[0] = Ops::fma2(a[0], s01, z01);
            a[1] = Ops::fma2(a[1], s01, z01);
            a[2] = Ops::fma2(a[2], s23, z23);
            a[3] = Ops::fma2(a[3], s23, z23);
            Ops::mma16816(O[2 * jj + 1], a, pf);
        }
    }
    else if constexpr (KVBITS == 4) {
        uint32_t p0 = lds32(st + Cfg::kOffVP + tokV0 * 4), p1 = lds32(st + Cfg::kOffVP + tokV1 * 4);
        uint32_t p2 = lds32(st + Cfg::kOffVP + tokV2 * 4), p3 = lds32(st + Cfg::kOffVP + tokV3 * 4);
        if (residue) {
            if (t0 + tokV0 >= k_len) p0 = 0;
            if (t0 + tokV1 >= k_len) p1 = 0;
            if (t0 + tokV2 >= k_len) p2 = 0;
            if (t0 + tokV3 >= k_len) p3 = 0;
        }
        const uint32_t s01 = __byte_perm(p0, p1, 0x5410), z01 = __byte_perm(p0, p1, 0x7632);
        const uint32_t s23 = __byte_perm(p2, p3, 0x5410), z23 = __byte_perm(p2, p3, 0x7632);
        const uint2    v0 = lds64(st + Cfg::kOffV + offV0 + 8 * g), v1 = lds64(st + Cfg::kOffV + offV1 + 8 * g);
        const uint2    v2 = lds64(st + Cfg::kOffV + offV2 + 8 * g), v3 = lds64(st + Cfg::kOffV + offV3 + 8 * g);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const uint32_t A = (&v0.x)[m], Bq = (&v1.x)[m], C = (&v2.x)[m], Dq = (&v3.x)[m];
            // lo16 of a word: elements e0,e2,e4,e6 (nibbles 0-3); hi16: e1,e3,e5,e7
            const uint32_t Ye = __byte_perm(A, Bq, 0x5410), Yo = __byte_perm(A, Bq, 0x7632);
            const uint32_t Ze = __byte_perm(C, Dq, 0x5410), Zo = __byte_perm(C, Dq, 0x7632);
            uint32_t       he[4], ho[4], ge[4], go[4];
            Ops::cvt_u4x8(Ye, he);  // he[x] = (A.e(2x), B.e(2x))   [(n0,n4),(n1,n5),..] of Ye
            Ops::cvt_u4x8(Yo, ho);  // ho[x] = (A.e(2x+1), B.e(2x+1))
            Ops::cvt_u4x8(Ze, ge);
            Ops::cvt_u4x8(Zo, go);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                // element index inside the word: even -> 2x (m-slot g), odd -> 2x+1 (m-slot g+8)
                uint32_t a[4];
                a[0] = Ops::fma2(he[x], s01, z01);
                a[1] = Ops::fma2(ho[x], s01, z01);
                a[2] = Ops::fma2(ge[x], s23, z23);
                a[3] = Ops::fma2(go[x], s23, z23);
                Ops::mma16816(O[4 * m + x], a, pf);
            }
        }
    }
    else {
        uint4 va[2], vb[2], vc[2], vd[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            va[i] = lds128(st + Cfg::kOffV + offV0 + 16 * (g + 8 * i));
            vb[i] = lds128(st + Cfg::kOffV + offV1 + 16 * (g + 8 * i));
            vc[i] = lds128(st + Cfg::kOffV + offV2 + 16 * (g + 8 * i));
            vd[i] = lds128(st + Cfg::kOffV + offV3 + 16 * (g + 8 * i));
        }
        if (residue) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            if (t0 + tokV0 >= k_len) va[0] = va[1] = z;
            if (t0 + tokV1 >= k_len) vb[0] = vb[1] = z;
            if (t0 + tokV2 >= k_len) vc[0] = vc[1] = z;
            if (t0 + tokV3 >= k_len) vd[0] = vd[1] = z;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t A = (&va[j >> 2].x)[j & 3], Bq = (&vb[j >> 2].x)[j & 3];
            const uint32_t C = (&vc[j >> 2].x)[j & 3], Dq = (&vd[j >> 2].x)[j & 3];
            uint32_t       a[4];
            a[0] = __byte_perm(A, Bq, 0x5410);  // (dim 2j, keys 2t,2t+1)
            a[1] = __byte_perm(A, Bq, 0x7632);  // (dim 2j+1, ...)
            a[2] = __byte_perm(C, Dq, 0x5410);
            a[3] = __byte_perm(C, Dq, 0x7632);
            Ops::mma16816(O[j], a, pf);
        }
    }

}

B200_DEVICE void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
B200_DEVICE void sts64(uint32_t addr, uint2 v)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// In-kernel prologue for the new token of (sequence b, kv-head h) at cache position ti (= tile slot `pos`): one warp,
// lane = 4 consecutive dims (the mapping of kv_prologue_kernel, whose arithmetic this repeats instruction for instruction).
template<class T, int KVBITS>
__device__ __noinline__ void new_token_kv(const KernelParams& p, int b, int h, int ti, int pos, int lane, uint32_t st)
{
    using Ops = TypeOps<T>;
    using Cfg = TileCfg<KVBITS>;
    const int ro_len = p.readonly_block_num ? p.readonly_block_num[b] * p.block_len : 0;
    if (ti < ro_len) {
        return;  // inside a shared read-only prefix block: the page already holds this position
    }
    const int tok = p.cu_q_len ? p.cu_q_len[b] : b;
    const int di  = lane * 4;
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
    float2 cs[2];
    rope_cs2(cs, p.rope, p.rope_table, tok, di, ti);
    rope_rotate4<T>(xk, cs, p.rope, di);

    constexpr int kRow          = Cfg::kRowBytes;
    const int     head_data_sz  = p.block_len * kRow;
    const int     head_param_sz = p.block_len * 4;
    char*         blk = p.block_ptrs[p.cu_block_nums[b] + ti / p.block_len] + p.layer_offset;
    const int     bt  = ti % p.block_len;
    char*         kd  = blk + (size_t)(h * 2) * head_data_sz + (size_t)bt * kRow;
    char*         vd  = kd + head_data_sz;
    const uint32_t srow = token_smem_offset<KVBITS>(pos);
    if constexpr (KVBITS == 16) {
        const uint2 uk = *reinterpret_cast<uint2*>(&xk[0]), uv = *reinterpret_cast<uint2*>(&xv[0]);
        *reinterpret_cast<uint2*>(kd + 8 * lane) = uk;
        *reinterpret_cast<uint2*>(vd + 8 * lane) = uv;
        sts64(st + Cfg::kOffK + srow + 8 * lane, uk);
        sts64(st + Cfg::kOffV + srow + 8 * lane, uv);
    }
    else {
        char* kp = blk + (size_t)p.num_kv_heads * 2 * head_data_sz + (size_t)(h * 2) * head_param_sz + (size_t)bt * 4;
        char* vp = kp + head_param_sz;
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
            uint32_t bytes, param;
            bool     valid;
            quantize_head<T, KVBITS>(kv ? xv : xk, lane, bytes, valid, param);
            char*          dd   = kv ? vd : kd;
            const uint32_t sdat = st + (kv ? Cfg::kOffV : Cfg::kOffK) + srow;
            if constexpr (KVBITS == 8) {
                *reinterpret_cast<uint32_t*>(dd + 4 * lane) = bytes;
                sts32(sdat + 4 * lane, bytes);
            }
            else if (valid) {
                *reinterpret_cast<uint32_t*>(dd + 4 * (lane >> 1)) = bytes;
                sts32(sdat + 4 * (lane >> 1), bytes);
            }
            if (lane == 0) {
                *reinterpret_cast<uint32_t*>(kv ? vp : kp) = param;
                sts32(st + (kv ? Cfg::kOffVP : Cfg::kOffKP) + pos * 4, param);
            }
        }
    }
}

B200_DEVICE SegCtas segment_ctas(const KernelParams& p, int b, int h, int grid_main)
{
    SegCtas         s;
    const int       Hkv   = p.num_kv_heads;
    const int       tiles = p.tile_prefix[b + 1] - p.tile_prefix[b];
    s.total               = (long long)p.tile_prefix[p.batch] * Hkv;
    s.s0                  = (long long)Hkv * p.tile_prefix[b] + (long long)h * tiles;
    const long long s1    = s.s0 + tiles;
    s.cf                  = (int)(((s.s0 + 1) * grid_main - 1) / s.total);
    s.cl                  = (int)((s1 * grid_main - 1) / s.total);
    s.first_slot          = (s.total * s.cf / grid_main != s.s0) ? 1 : 0;
    s.sparse              = s.total < grid_main;
    s.n                   = s.sparse ? tiles : s.cl - s.cf + 1;
    return s;
}

// Merge of the partials of one segment.  Same math as attention::reduce (reduce.cu:90-153): O = sum_k w_k O_k / sum_k w_k L_k,
// w_k = exp2((M_k - M) * c), pieces visited in ascending CTA order whoever runs the merge (deterministic).  Loads bypass L1
// (ld.global.cg): the partials may have been written by other SMs during this very kernel.
template<class T>
__device__ __noinline__ void merge_segment(const KernelParams& p, int b, int h, const SegCtas& sg, int grid_main, int tid,
                                           int nthreads)
{
    using Ops     = TypeOps<T>;
    const int   G = p.num_heads / p.num_kv_heads;
    const float c = p.qk_scale;
    auto slot_of  = [&](int k) {
        if (sg.sparse) {
            return 2 * (int)(((sg.s0 + k + 1) * grid_main - 1) / sg.total);  // one unit per CTA: always its first slot
        }
        const int cta = sg.cf + k;
        return 2 * cta + (k == 0 ? sg.first_slot : 0);
    };
    // thread -> 4 consecutive dims of one head (float4 partial loads)
    for (int i4 = tid; i4 < G * kHeadDim / 4; i4 += nthreads) {
        const int i  = i4 * 4;
        const int hd = i / kHeadDim;
        float     m  = -INFINITY;
        for (int k = 0; k < sg.n; ++k) {
            m = fmaxf(m, __ldcg(&p.partial_ML[((size_t)slot_of(k) * kMaxGroup + hd) * 2]));
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float  l   = 0.f;
        for (int k = 0; k < sg.n; ++k) {
            const int    slot = slot_of(k);
            const float2 ml   = __ldcg(reinterpret_cast<const float2*>(&p.partial_ML[((size_t)slot * kMaxGroup + hd) * 2]));
            const float  w    = (ml.x == -INFINITY) ? 0.f : exp2_approx((ml.x - m) * c);
            const float4 o    = __ldcg(reinterpret_cast<const float4*>(&p.partial_O[(size_t)slot * kMaxGroup * kHeadDim + i]));
            l += w * ml.y;
            acc.x += w * o.x, acc.y += w * o.y, acc.z += w * o.z, acc.w += w * o.w;
        }
        const int col = h * G * kHeadDim + i;
        T*        dst = reinterpret_cast<T*>(p.out)
                 + (p.out_mt > 0 ? ((size_t)(col >> 3) * p.out_mt + b) * 8 + (col & 7) : (size_t)b * p.out_stride + col);
        uint2 v;
        v.x = Ops::pack2(acc.x / l, acc.y / l);
        v.y = Ops::pack2(acc.z / l, acc.w / l);
        *reinterpret_cast<uint2*>(dst) = v;
    }
}

B200_DEVICE int atom_add_acq_rel_gpu(int* p, int v)
{
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

struct TileDesc {
    const char* blk;  // block base + layer offset
    int         h;
    int         bt;    // token offset of the tile inside the block
    int         rows;  // valid tokens in the tile
};

// Resident CTAs per SM the register budget must allow (the shared-memory ring is sized for the same numbers): the tile
// loop keeps ~125 registers live; the rare paths (new token's K/V, segment merge) are out-of-line calls so they cannot
// inflate it.  __grid_constant__: those calls take the parameter block by reference without a local copy.
template<int KVBITS>
constexpr int min_ctas_per_sm()
{
    return KVBITS == 16 ? 2 : 3;
}

template<class T, int KVBITS, int STAGES>
__global__ void __launch_bounds__(kThreads, min_ctas_per_sm<KVBITS>()) decode_attention_kernel(const __grid_constant__ KernelParams p)
{
    using Ops = TypeOps<T>;
    using Cfg = TileCfg<KVBITS>;
    using SS  = SharedStorage<T, KVBITS, STAGES>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SS&                                      ss = *reinterpret_cast<SS*>(smem_raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // q and the page of the new token come from the prologue kernel right before this one; the tile plan, the block table and
    // the lengths from kernels (or copies) several launches back.  p.early (the caller supplied the plan and the prologue kernel
    // triggers its dependents only AFTER its own griddepcontrol.wait, so everything older than the prologue is complete when
    // this grid starts): the producer warp does not wait here -- it fills the ring with tiles the prologue does not touch
    // (every tile but the last of a sequence) while the prologue runs, and waits before the first tile that holds a new token.
    pdl_launch_dependents();
    const bool early = p.early != 0 && warp == kConsumerWarps;
    if (!early) {
        pdl_wait();
    }

    const int       Hkv   = p.num_kv_heads;
    const int       G     = p.num_heads / Hkv;
    const long long total = (long long)p.tile_prefix[p.batch] * Hkv;
    const long long lo    = total * blockIdx.x / gridDim.x;
    const long long hi    = total * (blockIdx.x + 1) / gridDim.x;

    // the producer warp owns the ring's barriers: it initialises them and ARRIVES on named barrier 3; the consumers sync on it
    if (warp == kConsumerWarps) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(smem_u32(&ss.full[s]), 1);
                mbar_init(smem_u32(&ss.empty[s]), kConsumerWarps);
            }
            fence_mbar_init();
        }
        __syncwarp();
        named_bar_arrive(3, kThreads);
    }
    else {
        named_bar_sync(3, kThreads);
    }

    if (lo >= hi) {
        return;
    }

    const uint32_t full0  = smem_u32(&ss.full[0]);
    const uint32_t empty0 = smem_u32(&ss.empty[0]);
    const uint32_t stage0 = smem_u32(&ss.stage[0][0]);

    if (warp == kConsumerWarps) {
        // ================================ producer warp ================================
        // Tile descriptors (block pointer through the block table, valid rows) are fetched 32 at a time, one per
        // lane, one batch AHEAD of the copies being issued, so the dependent global loads
        // (cu_k_len -> cu_block_nums -> block_ptrs) never sit on the critical path of the TMA issue loop.
        const uint64_t policy         = make_evict_first_policy();
        const int      head_data_size = p.block_len * Cfg::kRowBytes;
        const int      head_param_sz  = p.block_len * 4;
        auto           fetch          = [&](long long base) {
            TileDesc        d{nullptr, 0, 0, 0};
            const long long idx = base + lane;
            if (idx < hi) {
                WorkIter it{p.batch, Hkv, p.tile_prefix};
                it.seek(idx);
                const int k_len = p.cu_k_len[it.b + 1] - p.cu_k_len[it.b];
                const int t0    = it.tile * kTile;
                d.rows          = min(kTile, k_len - t0) | (t0 + kTile >= k_len ? 128 : 0);  // bit 7: holds the new token
                d.h             = it.h;
                d.bt            = t0 % p.block_len;
                d.blk           = p.block_ptrs[p.cu_block_nums[it.b] + t0 / p.block_len] + p.layer_offset;
            }
            return d;
        };
        int      s  = 0;
        uint32_t ph = 0;
        bool     pending = early;  // griddepcontrol.wait still to come
        int      issued  = 0;
        TileDesc cur = fetch(lo);
        // Copy roles: lanes 0-7 move the four 16-row groups of K (even) and V (odd), lanes 8/9 the K/V quant params.
        // The ten copies of one tile are issued by ten lanes in ONE pass (a single lane issuing them serially costs
        // ~280 issue slots per tile on the scheduler this warp shares with a consumer warp).
        const int  my_gq    = (lane >> 1) & 3;
        const bool my_v     = lane & 1;
        const bool my_data  = lane < 8;
        const bool my_param = Cfg::kQuant && (lane == 8 || lane == 9);
        for (long long base = lo; base < hi; base += 32) {
            const TileDesc nxt = fetch(base + 32);
            const int      cnt = (int)min((long long)32, hi - base);
            for (int l = 0; l < cnt; ++l) {
                const unsigned long long blk = __shfl_sync(0xffffffffu, (unsigned long long)cur.blk, l);
                const int packed = __shfl_sync(0xffffffffu, cur.h | (cur.bt << 8) | (cur.rows << 24), l);
                const int h = packed & 0xff, bt = (packed >> 8) & 0xffff, rows = (packed >> 24) & 127;
                if (pending && (((packed >> 31) & 1) || issued == STAGES)) {
                    pdl_wait();  // warp-uniform: the prologue's stores (new token's K / V, their params) are visible from here
                    pending = false;
                }
                ++issued;
                const uint32_t bar = full0 + 8 * s;
                if (lane == 0) {
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    const int prow = (rows + 3) & ~3;
                    mbar_expect_tx(bar, 2 * rows * Cfg::kRowBytes + (Cfg::kQuant ? 2 * prow * 4 : 0));
                }
                __syncwarp();
                const uint32_t st = stage0 + s * (uint32_t)Cfg::kStageBytes;
                const char*    hb = reinterpret_cast<const char*>(blk);
                if (my_data) {
                    const int r = min(kGroupRows, rows - my_gq * kGroupRows);
                    if (r > 0) {
                        const char* src = hb + (size_t)(h * 2 + (my_v ? 1 : 0)) * head_data_size
                                          + (size_t)(bt + my_gq * kGroupRows) * Cfg::kRowBytes;
                        bulk_g2s_hint(st + (my_v ? Cfg::kOffV : Cfg::kOffK) + my_gq * Cfg::kGroupSize, src,
                                      r * Cfg::kRowBytes, bar, policy);
                    }
                }
                else if (my_param) {
                    const int   prow = (rows + 3) & ~3;
                    const char* src  = hb + (size_t)Hkv * 2 * head_data_size
                                      + (size_t)(h * 2 + (my_v ? 1 : 0)) * head_param_sz + (size_t)bt * 4;
                    bulk_g2s_hint(st + (my_v ? Cfg::kOffVP : Cfg::kOffKP), src, prow * 4, bar, policy);
                }
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
            cur = nxt;
        }
        return;
    }

    // ================================== consumers ==================================
    ConsumerCtx<T, KVBITS> cx;
    cx.warp = warp;
    cx.g    = lane >> 2;
    cx.t    = lane & 3;
    const int g = cx.g, t = cx.t;
    cx.tokA = slot_token(g, warp), cx.tokB = slot_token(g + 8, warp);
    cx.tokV0 = slot_token(2 * t, warp), cx.tokV1 = slot_token(2 * t + 1, warp);
    cx.tokV2 = slot_token(2 * t + 8, warp), cx.tokV3 = slot_token(2 * t + 9, warp);
    cx.offA = token_smem_offset<KVBITS>(cx.tokA), cx.offB = token_smem_offset<KVBITS>(cx.tokB);
    cx.offV0 = token_smem_offset<KVBITS>(cx.tokV0), cx.offV1 = token_smem_offset<KVBITS>(cx.tokV1);
    cx.offV2 = token_smem_offset<KVBITS>(cx.tokV2), cx.offV3 = token_smem_offset<KVBITS>(cx.tokV3);

    uint32_t qf[8][2];  // Q as B-operand fragments, one per k-step
    float    O[8][4];   // O^T accumulators: m-tile j -> {(slot g,head 2t),(g,2t+1),(g+8,2t),(g+8,2t+1)}
    float    M[2], L[2];

    const float c = p.qk_scale;

    WorkIter it{p.batch, Hkv, p.tile_prefix};
    it.seek(lo);
    int       s   = 0;
    uint32_t  ph  = 0;
    long long idx = lo;

    while (idx < hi) {
        // ---------------- one segment: tiles [first_tile, first_tile + ntile) of (it.b, it.h) ----------------
        const int first_tile = it.tile;
        const int ntile      = (int)min((long long)(it.tiles - it.tile), hi - idx);
        const int k_len      = p.cu_k_len[it.b + 1] - p.cu_k_len[it.b];
        {
            if (p.fused & 1) {
                // RoPE applied in registers: every lane rotates the 4-dim chunks of its own B fragments
                const int tokq = p.cu_q_len ? p.cu_q_len[it.b] : it.b;
                const T*  qb   = reinterpret_cast<const T*>(p.q) + (size_t)tokq * p.stride + (size_t)it.h * G * kHeadDim;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint2 v = make_uint2(0u, 0u);
                    if (g < G) {
                        const int di = qk_dim0<KVBITS>(t, j);
                        T         x[4];
                        load4(qb + (size_t)g * kHeadDim + di, x);
                        if (p.q_bias) {
                            T bq[4];
                            load4(reinterpret_cast<const T*>(p.q_bias) + (size_t)(it.h * G + g) * kHeadDim + di, bq);
#pragma unroll
                            for (int i = 0; i < 4; ++i) x[i] = Ops::add(x[i], bq[i]);
                        }
                        float2 cs[2];
                        rope_cs2(cs, p.rope, p.rope_table, tokq, di, k_len - 1);
                        rope_rotate4<T>(x, cs, p.rope, di);
                        v = *reinterpret_cast<uint2*>(&x[0]);
                    }
                    qf[j][0] = v.x;
                    qf[j][1] = v.y;
                }
            }
            else {
                const T* qb = reinterpret_cast<const T*>(p.q_rot) + ((size_t)it.b * p.num_heads + (size_t)it.h * G) * kHeadDim;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint2 v = make_uint2(0u, 0u);
                    if (g < G) {
                        v = *reinterpret_cast<const uint2*>(qb + (size_t)g * kHeadDim + qk_dim0<KVBITS>(t, j));
                    }
                    qf[j][0] = v.x;
                    qf[j][1] = v.y;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                O[j][0] = O[j][1] = O[j][2] = O[j][3] = 0.f;
            }
            M[0] = M[1] = -INFINITY;
            L[0] = L[1] = 0.f;
        }
        // The tile that holds the new token (position k_len - 1, always the sequence's last tile) is peeled out of the loop:
        // the out-of-line prologue call must not sit inside the loop (the ABI would pin the loop state in local memory).
        const bool peel  = (p.fused & 1) && first_tile + ntile == it.tiles;
        const int  nloop = ntile - (peel ? 1 : 0);
        for (int tt = 0; tt < nloop; ++tt) {
            const int  t0      = (first_tile + tt) * kTile;
            const bool residue = t0 + kTile > k_len;
            mbar_wait(full0 + 8 * s, ph);
            process_tile<T, KVBITS>(cx, stage0 + s * (uint32_t)Cfg::kStageBytes, residue, k_len, t0, c, qf, O, M, L);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty0 + 8 * s);
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        if (peel) {
            const int t0 = (it.tiles - 1) * kTile;
            mbar_wait(full0 + 8 * s, ph);
            // The new token's slot belongs to exactly one consumer warp (slot_token: token & 15 = 4 * warp + i): that warp runs
            // the prologue for this (sequence, kv-head) -- bias, RoPE(k), quantise, store into the page for later steps -- and
            // overwrites the row the bulk copy brought in (stale bytes) in ITS OWN view of the tile; no other warp reads it.
            const int pos = k_len - 1 - t0;
            if (((pos >> 2) & 3) == warp) {
                new_token_kv<T, KVBITS>(p, it.b, it.h, k_len - 1, pos, lane, stage0 + s * (uint32_t)Cfg::kStageBytes);
                __syncwarp();
            }
            process_tile<T, KVBITS>(cx, stage0 + s * (uint32_t)Cfg::kStageBytes, true, k_len, t0, c, qf, O, M, L);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty0 + 8 * s);
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        const bool last_tile = first_tile + ntile == it.tiles;

        // ---------------- segment end: merge the 4 consumer warps, write final or partial ----------------
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            L[q] += __shfl_xor_sync(0xffffffffu, L[q], 4);
            L[q] += __shfl_xor_sync(0xffffffffu, L[q], 8);
            L[q] += __shfl_xor_sync(0xffffffffu, L[q], 16);
        }
        if (g == 0) {
            ss.merge_M[warp][2 * t]     = M[0];
            ss.merge_M[warp][2 * t + 1] = M[1];
            ss.merge_L[warp][2 * t]     = L[0];
            ss.merge_L[warp][2 * t + 1] = L[1];
        }
        named_bar_sync(1, kConsumerWarps * 32);
        float sc[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int hd = 2 * t + q;
            float     m  = ss.merge_M[0][hd];
            for (int w = 1; w < kConsumerWarps; ++w) m = fmaxf(m, ss.merge_M[w][hd]);
            float l = 0.f;
            for (int w = 0; w < kConsumerWarps; ++w) {
                const float mw = ss.merge_M[w][hd];
                l += (mw == -INFINITY) ? 0.f : exp2_approx((mw - m) * c) * ss.merge_L[w][hd];
            }
            sc[q] = (M[q] == -INFINITY) ? 0.f : exp2_approx((M[q] - m) * c);
            if (warp == 0 && g == 0) {
                ss.final_M[hd] = m;
                ss.final_L[hd] = l;
            }
        }
        // fixed order (w0 + w2) + (w1 + w3)
        float* mo = &ss.merge_O[warp & 1][0][0];
        if (warp < 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = pv_dim<KVBITS>(g, j, 0);
                *reinterpret_cast<float2*>(mo + merge_idx(2 * t, d))     = make_float2(O[j][0] * sc[0], O[j][2] * sc[0]);
                *reinterpret_cast<float2*>(mo + merge_idx(2 * t + 1, d)) = make_float2(O[j][1] * sc[1], O[j][3] * sc[1]);
            }
        }
        named_bar_sync(1, kConsumerWarps * 32);
        if (warp >= 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d  = pv_dim<KVBITS>(g, j, 0);
                float2*   a0 = reinterpret_cast<float2*>(mo + merge_idx(2 * t, d));
                float2*   a1 = reinterpret_cast<float2*>(mo + merge_idx(2 * t + 1, d));
                float2    x0 = *a0, x1 = *a1;
                x0.x += O[j][0] * sc[0], x0.y += O[j][2] * sc[0];
                x1.x += O[j][1] * sc[1], x1.y += O[j][3] * sc[1];
                *a0 = x0, *a1 = x1;
            }
        }
        named_bar_sync(1, kConsumerWarps * 32);
        const bool complete = (first_tile == 0) && last_tile;
        const int  ctid     = threadIdx.x;  // 0..127
        const float* m0     = &ss.merge_O[0][0][0];
        const float* m1     = &ss.merge_O[1][0][0];
        if (complete) {
            for (int i = ctid; i < G * kHeadDim / 2; i += kConsumerWarps * 32) {
                const int    hd  = (2 * i) / kHeadDim;
                const int    d   = (2 * i) % kHeadDim;
                const float  inv = 1.f / ss.final_L[hd];
                const int    mi  = merge_idx(hd, d);
                const float2 a = *reinterpret_cast<const float2*>(m0 + mi), b2 = *reinterpret_cast<const float2*>(m1 + mi);
                const int    col = it.h * G * kHeadDim + 2 * i;
                T*           dst = reinterpret_cast<T*>(p.out)
                         + (p.out_mt > 0 ? ((size_t)(col >> 3) * p.out_mt + it.b) * 8 + (col & 7) :
                                           (size_t)it.b * p.out_stride + col);
                *reinterpret_cast<uint32_t*>(dst) = Ops::pack2((a.x + b2.x) * inv, (a.y + b2.y) * inv);
            }
        }
        else {
            const int slot = 2 * blockIdx.x + (idx == lo ? 0 : 1);
            float*    po   = p.partial_O + (size_t)slot * kMaxGroup * kHeadDim;
            for (int i = ctid; i < G * kHeadDim / 2; i += kConsumerWarps * 32) {
                const int    hd = (2 * i) / kHeadDim, d = (2 * i) % kHeadDim;
                const int    mi = merge_idx(hd, d);
                const float2 a = *reinterpret_cast<const float2*>(m0 + mi), b2 = *reinterpret_cast<const float2*>(m1 + mi);
                *reinterpret_cast<float2*>(po + 2 * i) = make_float2(a.x + b2.x, a.y + b2.y);
            }
            if (ctid < G) {
                p.partial_ML[((size_t)slot * kMaxGroup + ctid) * 2]     = ss.final_M[ctid];
                p.partial_ML[((size_t)slot * kMaxGroup + ctid) * 2 + 1] = ss.final_L[ctid];
            }
            if (p.fused & 2) {
                // In-kernel fix-up: arrive on the segment's counter (the CTA barrier orders every thread's partial stores
                // before thread 0's gpu-scope acq_rel atomic); the CTA that arrives LAST merges all partials of the segment
                // in CTA order -- the same sum whoever merges -- and re-arms the counter.
                named_bar_sync(1, kConsumerWarps * 32);
                if (ctid == 0) {
                    ss.seg        = segment_ctas(p, it.b, it.h, p.grid_main);
                    int*      ctr = p.seg_counters + it.b * Hkv + it.h;
                    const int old = atom_add_acq_rel_gpu(ctr, 1);
                    ss.fix_last   = (old == ss.seg.n - 1);
                    if (old == ss.seg.n - 1) {
                        *ctr = 0;  // every piece of the segment has arrived: nobody touches the counter again in this launch
                    }
                }
                named_bar_sync(1, kConsumerWarps * 32);
                if (ss.fix_last) {
                    merge_segment<T>(p, it.b, it.h, ss.seg, p.grid_main, ctid, kConsumerWarps * 32);
                }
            }
        }
        named_bar_sync(1, kConsumerWarps * 32);

        // advance to the next segment
        idx += ntile;
        it.tile += ntile - 1;
        it.next();
    }
}

// Fix-up reduction as a separate kernel (used when the in-kernel fix-up is off).
template<class T>
__global__ void decode_reduce_kernel(const KernelParams p, int grid_main)
{
    pdl_launch_dependents();
    pdl_wait();
    const int b     = blockIdx.y;
    const int h     = blockIdx.x;
    const int tiles = p.tile_prefix[b + 1] - p.tile_prefix[b];
    if (tiles == 0) {
        return;
    }
    __shared__ SegCtas s_seg;
    if (threadIdx.x == 0) {
        s_seg = segment_ctas(p, b, h, grid_main);
    }
    __syncthreads();
    if (s_seg.n <= 1) {
        return;  // written directly by the main kernel
    }
    merge_segment<T>(p, b, h, s_seg, grid_main, threadIdx.x, blockDim.x);
}

}  // namespace attn
}  // namespace b200
