// Paged-KV causal prefill ("context") attention for sm_100a -- SURVEY section 8f rank 1.
//
// Stands where the reference runs, for a chunk of q_len > 1 new tokens per sequence,
//   invokeProcessKV_v2  (kv_cache_utils_v2.cu:18-210)   -> b200_process_kv's prologue kernel (shared with decode)
//   invokeFlattenKV_v2  (kv_cache_utils_v2.cu:212-420)  -> fused here: pages are de-quantised into the operand tiles
//   dispatchAttention   (attention.cu, attention_universal.h:355-553 with the sm80 prefill mainloop)
// Not a port: S = Q.K^T and O += P.V run on tcgen05 tensor cores with fp32 accumulators in tensor memory.
//
// One CTA = one kv-head x one tile of UMMA-M = 128 query rows.  Rows stack the G q-heads of the kv-head
// (row = token_in_tile * G + head_in_group, 128 / G tokens per tile), so every K/V tile fetched and de-quantised from
// the pages serves all heads that share it.  Per 64-key tile, in phases separated by block barriers (two CTAs per SM
// overlap each other's phases; a warp-specialised pipeline inside one CTA is the planned next step):
//   A  128 threads gather the tile's K and V rows from the pages through the block table, de-quantise int8/int4 with
//      the decode kernel's arithmetic (exact int -> T, one T fma with the token's (scale, zero)) and store them as UMMA
//      K-major operand tiles: K as [d/8][key][8], V transposed as [key/8][d][8];
//   B  one thread issues 8 x tcgen05.mma (128x64x16, A = Q tile in smem, B = K tile) -> S in TMEM;
//   C  thread r owns row r: tcgen05.ld its 64 scores, causal mask, base-2 online softmax (same formula as the decode
//      kernel), P cast to T and written back to TMEM as the A operand of the next MMA;
//   D  one thread issues 4 x tcgen05.mma (128x128x16, A = P in TMEM, B = V^T tile) -> O_tile in TMEM;
//   E  thread r folds O_tile into its fp32 row accumulator with the rescale factor of phase C.
// TMEM: S 64 columns, P 32, O_tile 128 -> 256 allocated; 2 CTAs per SM fill the 512.
#pragma once

#include "common.cuh"
#include "umma.cuh"

namespace b200 {
namespace attn {

struct PrefillParams {
    const void* q_rot;  // [token][Hq][D], bias + RoPE applied (kv_prologue_kernel)
    void*       out;    // [token][Hq*D]
    int64_t     out_stride;
    char**      block_ptrs;
    const int*  cu_block_nums;
    const int*  cu_q_len;
    const int*  cu_k_len;
    int         layer_offset;
    int         block_len;
    int         batch;
    int         num_heads;
    int         num_kv_heads;
    float       qk_scale;  // log2(e) / sqrt(D)
};

constexpr int kPfRows    = 128;  // UMMA M: query rows per CTA
constexpr int kPfKeys    = 64;   // keys per tile: UMMA N of S, K extent of P.V
constexpr int kPfThreads = 128;
constexpr int kPfD       = 128;
constexpr int kPfQBytes  = kPfRows * kPfD * 2;
constexpr int kPfKBytes  = kPfKeys * kPfD * 2;
// V^T tile [key/8][d/8][8 d][8 keys]: a thread writes one whole 128-byte core matrix (8 chunks), so with the canonical
// 128-byte stride between core matrices every 16-byte store of a warp lands on the same four banks (8-way conflict).
// kPfVPad = 16 staggers successive core matrices by one bank group (SBO = 144) and makes the stores conflict-free --
// to be switched on once tools/umma_probe.cu (variants 4/5) has shown the tensor core accepts such an SBO.
constexpr int kPfVPad    = 0;
constexpr int kPfVSbo    = 128 + kPfVPad;          // byte stride between 8-dim groups of one key group
constexpr int kPfVLbo    = (kPfD / 8) * kPfVSbo;   // byte stride between 8-key groups
constexpr int kPfVBytes  = (kPfKeys / 8) * kPfVLbo;
// The second K/V buffer is reserved for the pipelined version; it also pins residency at two CTAs per SM, which is
// what the 2 x 256 TMEM columns allow.
constexpr int kPfSmem    = kPfQBytes + 2 * (kPfKBytes + kPfVBytes);
constexpr int kPfColS = 0, kPfColP = 64, kPfColO = 128, kPfTmemCols = 256;

// de-quantise 8 consecutive dims of one token into 4 packed T2 registers (dims (0,1) (2,3) (4,5) (6,7))
template<class T, int KVBITS>
B200_DEVICE void pf_load8(const char* row, int dim0, uint32_t s2, uint32_t z2, uint32_t (&r)[4])
{
    using Ops = TypeOps<T>;
    if constexpr (KVBITS == 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(row + dim0 * 2);
        r[0] = v.x, r[1] = v.y, r[2] = v.z, r[3] = v.w;
    }
    else if constexpr (KVBITS == 8) {
        const uint2 v = *reinterpret_cast<const uint2*>(row + dim0);
        Ops::cvt_u8x4(v.x, r[0], r[1]);
        Ops::cvt_u8x4(v.y, r[2], r[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = Ops::fma2(r[i], s2, z2);
    }
    else {
        // one u32 = 8 dims, nibble j holds element [0,2,4,6,1,3,5,7][j]; cvt_u4x8 pairs nibbles (j, j+4) = dims (2j, 2j+1)
        const uint32_t v = *reinterpret_cast<const uint32_t*>(row + dim0 / 2);
        Ops::cvt_u4x8(v, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = Ops::fma2(r[i], s2, z2);
    }
}

template<class T, int KVBITS>
__global__ void __launch_bounds__(kPfThreads) prefill_attention_kernel(const PrefillParams p)
{
    using Ops = TypeOps<T>;
    extern __shared__ __align__(1024) uint8_t pf_smem[];
    __shared__ uint64_t bar_s, bar_o;
    __shared__ uint32_t s_tmem;
    __shared__ int      s_b, s_q0;

    pdl_launch_dependents();
    pdl_wait();  // q_rot and the chunk's K/V pages come from the prologue kernel

    const int tid = threadIdx.x, warp = tid >> 5;
    const int G   = p.num_heads / p.num_kv_heads;
    const int TQ  = kPfRows / G;  // query tokens per tile
    const int h   = blockIdx.y;

    // ---- which (sequence, query tile) is this CTA ----
    if (tid == 0) {
        int x = blockIdx.x, b = 0;
        for (; b < p.batch; ++b) {
            const int nt = (p.cu_q_len[b + 1] - p.cu_q_len[b] + TQ - 1) / TQ;
            if (x < nt) {
                break;
            }
            x -= nt;
        }
        s_b  = b < p.batch ? b : -1;
        s_q0 = x * TQ;
        mbar_init(smem_u32(&bar_s), 1);
        mbar_init(smem_u32(&bar_o), 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int b = s_b;
    if (b < 0) {
        return;  // grid is an upper bound of the tile count
    }
    const int q0    = s_q0;
    const int q_len = p.cu_q_len[b + 1] - p.cu_q_len[b];
    const int k_len = p.cu_k_len[b + 1] - p.cu_k_len[b];
    const int hist  = k_len - q_len;
    const int nq    = min(TQ, q_len - q0);         // valid query tokens of this tile
    const int last_pos = hist + q0 + nq - 1;       // newest key any row of this tile may see
    const int n_tiles  = last_pos / kPfKeys + 1;

    // row of this thread
    const int  r     = tid;
    const int  tq    = r / G, g = r - tq * G;
    const bool valid = tq < nq;  // rows >= TQ * G (G not a power of two) have tq >= TQ >= nq
    const int  pos_q = valid ? hist + q0 + tq : 0x7fffffff;  // padding rows see everything (finite softmax, never stored)
    const size_t tok = (size_t)p.cu_q_len[b] + q0 + tq;

    const uint32_t q_op = smem_u32(pf_smem);
    const uint32_t k_op = q_op + kPfQBytes;
    const uint32_t v_op = k_op + kPfKBytes;

    // ---- Q tile -> smem [d/8][row][8] ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.q_rot)
                                                          + (tok * p.num_heads + (size_t)h * G + g) * kPfD);
#pragma unroll
        for (int c = 0; c < kPfD / 8; ++c) {
            const uint4 v = valid ? src[c] : make_uint4(0, 0, 0, 0);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q_op + (c * kPfRows + r) * 16), "r"(v.x), "r"(v.y),
                         "r"(v.z), "r"(v.w)
                         : "memory");
        }
    }
    if (warp == 0) {
        umma::tmem_alloc(smem_u32(&s_tmem), kPfTmemCols);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);

    constexpr bool kQuant        = KVBITS != 16;
    constexpr int  kRow          = kPfD * KVBITS / 8;
    const int      head_data_sz  = p.block_len * kRow;
    const int      head_param_sz = p.block_len * 4;
    const uint32_t idesc_s = umma::make_idesc_f16(kPfRows, kPfKeys, std::is_same<T, nv_bfloat16>::value);
    const uint32_t idesc_o = umma::make_idesc_f16(kPfRows, kPfD, std::is_same<T, nv_bfloat16>::value);
    const float    c       = p.qk_scale;

    float o[kPfD];
#pragma unroll
    for (int i = 0; i < kPfD; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        // ============================ A: pages -> operand tiles ============================
        {
            const int   key0 = j * kPfKeys;
            const char* blk  = p.block_ptrs[p.cu_block_nums[b] + key0 / p.block_len] + p.layer_offset;
            const int   bt0  = key0 % p.block_len;
            const char* kd   = blk + (size_t)(h * 2) * head_data_sz + (size_t)bt0 * kRow;
            const char* vd   = kd + head_data_sz;
            const char* kp   = blk + (size_t)p.num_kv_heads * 2 * head_data_sz + (size_t)(h * 2) * head_param_sz + (size_t)bt0 * 4;
            const char* vp   = kp + head_param_sz;
            // K: thread = (key = tid & 63, half = tid >> 6): 64 dims = 8 chunks, stored at [(d/8)][key]
            {
                const int  key  = tid & 63, half = tid >> 6;
                const bool live = key0 + key <= last_pos;  // rows past the newest visible key may be uninitialised
                uint32_t   s2 = 0, z2 = 0;
                if (kQuant && live) {
                    const uint32_t st = *reinterpret_cast<const uint32_t*>(kp + key * 4);
                    s2 = __byte_perm(st, st, 0x1010), z2 = __byte_perm(st, st, 0x3232);
                }
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint32_t x[4] = {0, 0, 0, 0};
                    if (live) {
                        pf_load8<T, KVBITS>(kd + (size_t)key * kRow, half * 64 + cc * 8, s2, z2, x);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(k_op + ((half * 8 + cc) * kPfKeys + key) * 16),
                                 "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3])
                                 : "memory");
                }
            }
            // V: thread = (key group kg = tid >> 4, dim group dg = tid & 15): an 8 x 8 block, transposed in registers
            {
                const int kg = tid >> 4, dg = tid & 15;
                uint32_t  a[8][4];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int  key  = kg * 8 + jj;
                    const bool live = key0 + key <= last_pos;
                    a[jj][0] = a[jj][1] = a[jj][2] = a[jj][3] = 0;  // P = 0 times garbage must stay 0
                    if (live) {
                        uint32_t s2 = 0, z2 = 0;
                        if (kQuant) {
                            const uint32_t st = *reinterpret_cast<const uint32_t*>(vp + key * 4);
                            s2 = __byte_perm(st, st, 0x1010), z2 = __byte_perm(st, st, 0x3232);
                        }
                        pf_load8<T, KVBITS>(vd + (size_t)key * kRow, dg * 8, s2, z2, a[jj]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // chunk of dim dg*8+i: keys kg*8 .. kg*8+7
                    const uint32_t sel = (i & 1) ? 0x7632 : 0x5410;
                    uint32_t       y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) y[q] = __byte_perm(a[2 * q][i >> 1], a[2 * q + 1][i >> 1], sel);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(v_op + kg * kPfVLbo + dg * kPfVSbo + i * 16), "r"(y[0]),
                                 "r"(y[1]), "r"(y[2]), "r"(y[3])
                                 : "memory");
                }
            }
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        umma::fence_before_sync();
        __syncthreads();
        // ============================ B: S = Q.K^T ============================
        if (tid == 0) {
            umma::fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < kPfD / 16; ++ks) {
                const uint64_t ad = umma::make_smem_desc(q_op + ks * 2 * (kPfRows * 16), kPfRows * 16, 128);
                const uint64_t bd = umma::make_smem_desc(k_op + ks * 2 * (kPfKeys * 16), kPfKeys * 16, 128);
                umma::mma_f16_ss(tmem + kPfColS, ad, bd, idesc_s, ks > 0);
            }
            umma::commit(smem_u32(&bar_s));
        }
        mbar_wait(smem_u32(&bar_s), ph);
        umma::fence_after_sync();
        // ============================ C: online softmax of row r ============================
        float alpha;
        {
            float mx = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < kPfKeys / 16; ++ch) {
                uint32_t v[16];
                umma::ld_32x32b_x16(trow + kPfColS + ch * 16, v);
                umma::wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float s = (j * kPfKeys + ch * 16 + i <= pos_q) ? __uint_as_float(v[i]) : -INFINITY;
                    mx            = fmaxf(mx, s);
                }
            }
            const float m_new = fmaxf(m, mx);  // finite: key 0 is visible to every row in tile 0
            alpha             = exp2_approx((m - m_new) * c);
            const float mc    = m_new * c;
            float       lsum  = 0.f;
            uint32_t    pk[kPfKeys / 2];
#pragma unroll
            for (int ch = 0; ch < kPfKeys / 16; ++ch) {
                uint32_t v[16];
                umma::ld_32x32b_x16(trow + kPfColS + ch * 16, v);
                umma::wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const int   k0 = j * kPfKeys + ch * 16 + i;
                    const float s0 = (k0 <= pos_q) ? __uint_as_float(v[i]) : -INFINITY;
                    const float s1 = (k0 + 1 <= pos_q) ? __uint_as_float(v[i + 1]) : -INFINITY;
                    const float p0 = exp2_approx(fmaf(s0, c, -mc));
                    const float p1 = exp2_approx(fmaf(s1, c, -mc));
                    lsum += p0 + p1;
                    pk[ch * 8 + i / 2] = Ops::pack2(p0, p1);
                }
            }
            l = l * alpha + lsum;
            m = m_new;
            // P (T) -> TMEM as the A operand: column c holds keys (2c, 2c+1) of this row
            {
                uint32_t lo[16], hi[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) lo[i] = pk[i], hi[i] = pk[16 + i];
                umma::st_32x32b_x16(trow + kPfColP, lo);
                umma::st_32x32b_x16(trow + kPfColP + 16, hi);
                umma::wait_st();
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        // ============================ D: O_tile = P.V ============================
        if (tid == 0) {
            umma::fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < kPfKeys / 16; ++ks) {
                const uint64_t bd = umma::make_smem_desc(v_op + ks * 2 * kPfVLbo, kPfVLbo, kPfVSbo);
                umma::mma_f16_ts(tmem + kPfColO, tmem + kPfColP + ks * 8, bd, idesc_o, ks > 0);
            }
            umma::commit(smem_u32(&bar_o));
        }
        mbar_wait(smem_u32(&bar_o), ph);
        umma::fence_after_sync();
        // ============================ E: O = O * alpha + O_tile ============================
#pragma unroll
        for (int ch = 0; ch < kPfD / 16; ++ch) {
            uint32_t v[16];
            umma::ld_32x32b_x16(trow + kPfColO + ch * 16, v);
            umma::wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[ch * 16 + i] = fmaf(o[ch * 16 + i], alpha, __uint_as_float(v[i]));
        }
        // the next tile's phase A overwrites K/V (their MMAs completed: bar_o) and its phase B overwrites S after a
        // block barrier that follows these loads
    }

    // ---- out = O / L ----
    if (valid) {
        const float inv = 1.f / l;
        uint4*      dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.out) + tok * p.out_stride + ((size_t)h * G + g) * kPfD);
#pragma unroll
        for (int cc = 0; cc < kPfD / 8; ++cc) {
            uint4 v;
            v.x = Ops::pack2(o[cc * 8 + 0] * inv, o[cc * 8 + 1] * inv);
            v.y = Ops::pack2(o[cc * 8 + 2] * inv, o[cc * 8 + 3] * inv);
            v.z = Ops::pack2(o[cc * 8 + 4] * inv, o[cc * 8 + 5] * inv);
            v.w = Ops::pack2(o[cc * 8 + 6] * inv, o[cc * 8 + 7] * inv);
            dst[cc] = v;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::tmem_dealloc(tmem, kPfTmemCols);
    }
}

}  // namespace attn
}  // namespace b200
