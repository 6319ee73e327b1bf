"""CPU oracle: TurboMind paged-KV decode attention (test infrastructure only).

Restates, in numpy, the algorithm of
  src/turbomind/kernels/attention/block.h:124-219           (KV page layout)
  src/turbomind/kernels/attention/quantization.h:341-366    (warp_stats: scale/zero)
  src/turbomind/kernels/attention/quantization.h:427-490    (ConvertKvCache<T,u8/u4>, pack)
  src/turbomind/kernels/attention/quantization.h:491-524,663-704 (u4/u8 -> T)
  src/turbomind/kernels/attention/impl_81616.h:338-348,463-474   (__hfma(x, scale, zero))
  src/turbomind/kernels/attention/rotary_embedding.h:10-52,168-181 (FastRoPE)
  src/turbomind/models/attention_weight.cc:37-93            (rope kernel params)
  src/turbomind/kernels/attention/attention_universal.h:168-348  (Prologue)
  src/turbomind/kernels/attention/impl_81616.h:510-617      (Softmax / ConvertStoP)
  src/turbomind/kernels/attention/reduce.cu:90-153          (split-K merge)

Numerics contract: T-typed ops are single RNE-rounded ops; where the
reference's release build uses hardware approximations (ex2.approx / sin.approx /
div.approx under --use_fast_math) the oracle -- and our CUDA kernels -- use the
correctly rounded operation instead.
"""
from dataclasses import dataclass, field
import numpy as np

from .numerics import F16, BF16, rn, to_bits, from_bits, t_add, t_sub, t_mul, t_fma, f32

ROPE_NULL, ROPE_DEFAULT, ROPE_LINEAR, ROPE_DYNAMIC, ROPE_YARN, ROPE_LLAMA3 = range(6)


# --------------------------------------------------------------------------------------
# page layout  (block.h:124-219)
# --------------------------------------------------------------------------------------
@dataclass
class KvLayout:
    """Byte layout of one layer inside one KV block: L(H 2 S D q) + (H 2 S 2 T)."""
    head_num: int          # local kv heads
    block_len: int
    head_dim: int
    q_bits: int            # 16 (T), 8 or 4
    t_bits: int = 16

    @property
    def quant(self):
        return self.q_bits != 16

    def token_data_size(self):
        return self.q_bits * self.head_dim // 8

    def token_param_size(self):
        return (self.t_bits * 2 // 8) if self.quant else 0

    def head_data_size(self):
        return self.block_len * self.token_data_size()

    def head_param_size(self):
        return self.block_len * self.token_param_size()

    def layer_size(self):
        return self.head_num * 2 * (self.head_data_size() + self.head_param_size())

    def k_data(self, head, ti):
        return head * 2 * self.head_data_size() + ti * self.token_data_size()

    def v_data(self, head, ti):
        return self.k_data(head, ti) + self.head_data_size()

    def k_param(self, head, ti):
        return (self.head_num * 2 * self.head_data_size() + head * 2 * self.head_param_size()
                + ti * self.token_param_size())

    def v_param(self, head, ti):
        return self.k_param(head, ti) + self.head_param_size()


# --------------------------------------------------------------------------------------
# RoPE  (attention_weight.cc:37-93, rotary_embedding.h)
# --------------------------------------------------------------------------------------
@dataclass
class RopeParam:
    type: int = ROPE_DEFAULT
    dim: int = 128
    base: float = 10000.0
    factor: float = 0.0                  # linear scaling factor (0 -> 1)
    max_position_embeddings: int = 0
    llama3_low_freq_factor: float = 1.0
    llama3_high_freq_factor: float = 4.0
    llama3_original_max_position_embeddings: int = 8192
    yarn_attention_factor: float = 1.0
    yarn_beta_fast: float = 32.0
    yarn_beta_slow: float = 1.0

    # derived kernel params (init_rope_kernel_param)
    def kernel_params(self):
        kp = {}
        kp["type"] = self.type
        kp["dim"] = self.dim
        # -std::log2(rope.base) / rope.dim  (float log2, float divide)
        kp["scale_factor"] = np.float32(-np.log2(np.float32(self.base)) / np.float32(self.dim))
        if self.type == ROPE_DYNAMIC:
            kp["inv_factor"] = np.float32(1.0)
        else:
            kp["inv_factor"] = np.float32(1.0 / self.factor) if self.factor != 0.0 else np.float32(1.0)
        kp["attention_scaling"] = np.float32(1.0)
        if self.type == ROPE_LLAMA3:
            inv_diff = np.float32(1.0 / (self.llama3_high_freq_factor - self.llama3_low_freq_factor))
            kp["llama3_alpha"] = np.float32(
                self.llama3_original_max_position_embeddings / (2 * 3.14159265358979323846) * inv_diff)
            kp["llama3_beta"] = np.float32(np.float32(self.llama3_low_freq_factor) * inv_diff)
        if self.type == ROPE_YARN:
            PI = 3.14159265358979323846

            def corr_dim(num_rot):
                return (self.dim * np.log(self.max_position_embeddings / (num_rot * 2 * PI))) / (2 * np.log(self.base))
            low = np.float32(max(np.floor(corr_dim(self.yarn_beta_fast)), 0.0))
            high = np.float32(min(np.ceil(corr_dim(self.yarn_beta_slow)), self.dim - 1.0))
            if low == high:
                high = np.float32(high + np.float32(0.001))
            kp["yarn_ramp_inv_factor_div_2"] = np.float32(1.0 / float(high - low) / 2.0)
            kp["yarn_ramp_inv_factor_mul_min"] = np.float32(1.0 / float(high - low) * float(low))
            kp["attention_scaling"] = np.float32(self.yarn_attention_factor)
        return kp


def rope_inv_freq(rope: RopeParam):
    """inv_freq for every even dim index i (pairs (i, i+1)); float32.  rotary_embedding.h:10-52."""
    kp = rope.kernel_params()
    D = rope.dim
    idx = np.arange(0, D, 2, dtype=np.int32)
    # (idx + i) * scale_factor : int->float convert, f32 multiply
    e = f32(idx.astype(np.float32) * kp["scale_factor"])
    # exp2f: correctly rounded (via float64)
    freq = f32(np.exp2(e.astype(np.float64)))
    if rope.type in (ROPE_DEFAULT, ROPE_LINEAR, ROPE_DYNAMIC):
        return f32(kp["inv_factor"] * freq)
    if rope.type == ROPE_LLAMA3:
        # smooth = clamp(fma(alpha, freq, -beta), 0, 1)
        sm = f32(kp["llama3_alpha"].astype(np.float64) * freq.astype(np.float64) - np.float64(kp["llama3_beta"]))
        sm = np.clip(sm, np.float32(0), np.float32(1))
        a = f32(f32(np.float32(1) - sm) * freq)
        b = f32(a * kp["inv_factor"])
        c = f32(sm * freq)
        return f32(b + c)
    if rope.type == ROPE_YARN:
        alpha = f32(f32(idx.astype(np.float32) * kp["yarn_ramp_inv_factor_div_2"]) - kp["yarn_ramp_inv_factor_mul_min"])
        alpha = np.clip(alpha, np.float32(0), np.float32(1))
        t = f32(f32(freq * alpha) * f32(np.float32(1) - kp["inv_factor"]))
        return f32(freq - t)
    raise ValueError(rope.type)


def apply_rope(x, pos, rope: RopeParam, dtype):
    """x: [..., D] float32 (exact in T); pos: int array broadcastable to x[..., 0].
    Adjacent-pair rotation computed in T (rotary_embedding.h:168-181).  The reference binary contracts each
    `(T)c*x0 - (T)s*x1` into HMUL2 + HFMA2 (mul.f16 / sub.f16 carry no .rn, so ptxas may fuse; checked in the SASS of
    oracle/_ref: decoding_sm80_128 / ProcessKV_v2):  x0' = fma(c, x0, -RN(s*x1)) ; x1' = fma(c, x1, RN(s*x0))."""
    if rope.type == ROPE_NULL:
        return f32(x).copy()
    kp = rope.kernel_params()
    inv_freq = rope_inv_freq(rope)                                  # [D/2]
    pos = np.asarray(pos)
    ang = f32(pos.astype(np.float32)[..., None] * inv_freq)         # f32 mul
    c = f32(np.cos(ang.astype(np.float64)))
    s = f32(np.sin(ang.astype(np.float64)))
    s = f32(s * kp["attention_scaling"])
    c = f32(c * kp["attention_scaling"])
    cT, sT = rn(c, dtype), rn(s, dtype)
    x = f32(x)
    nd = rope.dim
    x0, x1 = x[..., 0:nd:2], x[..., 1:nd:2]
    y0 = t_fma(cT, x0, -t_mul(sT, x1, dtype), dtype)
    y1 = t_fma(cT, x1, t_mul(sT, x0, dtype), dtype)
    out = x.copy()
    out[..., 0:nd:2] = y0
    out[..., 1:nd:2] = y1
    return out


# --------------------------------------------------------------------------------------
# KV quantisation  (quantization.h)
# --------------------------------------------------------------------------------------
def kv_quant_params(x, bits, dtype):
    """x [..., D] (exact in T) -> (scale, zero) in T.   warp_stats, quantization.h:341-366."""
    x = f32(x)
    mn = x.min(axis=-1)
    mx = x.max(axis=-1)
    inv_q_max = np.float32(1.0) / np.float32((1 << bits) - 1)
    scale = f32(f32(mx - mn) * inv_q_max)
    return rn(scale, dtype), mn.copy()


def kv_quantize(x, scale, zero, bits, dtype):
    """ConvertKvCache<T,u8|u4>: q = cvt.rni.sat.u8.T((x - zero) * inv_scale), u4: min(q, 15).
    inv_scale = (T)(1.f / (float)scale).  quantization.h:427-447, 449-490, 213-255."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv_scale = rn(np.float32(1.0) / f32(scale), dtype)
        y = t_mul(t_sub(x, zero[..., None], dtype), inv_scale[..., None], dtype)
    q = np.rint(y)                       # RNE
    q = np.where(np.isnan(q), 0.0, q)    # cvt.sat: NaN -> 0
    q = np.clip(q, 0, 255)
    if bits == 4:
        q = np.minimum(q, 15)
    return q.astype(np.uint8)


_U4_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]   # element index stored in nibble j of each u32 (pack(), :461-474)


def pack_u4(q):
    """q [..., D] uint8 (0..15) -> bytes [..., D/2] in the reference's nibble order:
    nibble j of each 32-bit word holds element _U4_ORDER[j] of the 8-element group."""
    q = np.asarray(q, np.uint8)
    g = q.reshape(q.shape[:-1] + (-1, 8)).astype(np.uint32)
    w = np.zeros(g.shape[:-1], np.uint32)
    for j, e in enumerate(_U4_ORDER):
        w |= g[..., e] << np.uint32(4 * j)
    return w.view(np.uint8).reshape(q.shape[:-1] + (-1,)) if w.flags.c_contiguous else \
        np.ascontiguousarray(w).view(np.uint8).reshape(q.shape[:-1] + (-1,))


def unpack_u4(b, D):
    """inverse of pack_u4: bytes [..., D/2] -> uint8 [..., D].  cvt_f16x8_u4 ordering (:491-524)."""
    b = np.ascontiguousarray(b, np.uint8)
    w = b.view(np.uint32).reshape(b.shape[:-1] + (D // 8,))
    out = np.zeros(b.shape[:-1] + (D // 8, 8), np.uint8)
    for j, e in enumerate(_U4_ORDER):
        out[..., e] = (w >> np.uint32(4 * j)) & 0xF
    return out.reshape(b.shape[:-1] + (D,))


def kv_dequant(q, scale, zero, dtype):
    """x = __hfma((T)q, scale, zero)   (impl_81616.h:338-348)."""
    return t_fma(q.astype(np.float32), f32(scale)[..., None], f32(zero)[..., None], dtype)


# --------------------------------------------------------------------------------------
# paged cache container
# --------------------------------------------------------------------------------------
class PagedKvCache:
    """A pool of KV blocks (one bytearray per block, all layers inside one block) +
    per-sequence block tables, exactly the structure BlockIteratorParams describes
    (attention_params.h:20-25; unified_attention_layer.cc:163-171 for layer offsets)."""

    def __init__(self, num_blocks, layers, layout: KvLayout, dtype=F16, seed=0):
        self.layout = layout
        self.layers = layers
        self.dtype = dtype
        self.block_bytes = layers * layout.layer_size()
        self.pool = np.zeros((num_blocks, self.block_bytes), np.uint8)
        self.num_blocks = num_blocks

    def layer_offset(self, layer):
        return layer * self.layout.layer_size()

    def write_token(self, block_id, layer, head, ti, k, v):
        """k, v: [D] float32 exact in T (already biased + rotated).  Quantise + store."""
        L, off = self.layout, self.layer_offset(layer)
        blk = self.pool[block_id]
        if L.quant:
            for x, d_off, p_off in ((k, L.k_data(head, ti), L.k_param(head, ti)),
                                    (v, L.v_data(head, ti), L.v_param(head, ti))):
                scale, zero = kv_quant_params(x[None], L.q_bits, self.dtype)
                q = kv_quantize(x[None], scale, zero, L.q_bits, self.dtype)[0]
                data = q if L.q_bits == 8 else pack_u4(q)
                blk[off + d_off: off + d_off + data.size] = data
                blk[off + p_off: off + p_off + 4] = to_bits(np.array([scale[0], zero[0]], np.float32),
                                                            self.dtype).view(np.uint8)
        else:
            for x, d_off in ((k, L.k_data(head, ti)), (v, L.v_data(head, ti))):
                blk[off + d_off: off + d_off + 2 * x.size] = to_bits(x, self.dtype).view(np.uint8)

    def write_tokens(self, block_ids, layer, head, start, k, v):
        """Vectorised: k, v [S, D]; sequence positions start..start+S-1 through block table block_ids."""
        L, off = self.layout, self.layer_offset(layer)
        S = k.shape[0]
        pos = start + np.arange(S)
        bi = np.asarray(block_ids)[pos // L.block_len]
        ti = pos % L.block_len
        for x, is_v in ((k, 0), (v, 1)):
            d_base = off + head * 2 * L.head_data_size() + is_v * L.head_data_size()
            if L.quant:
                scale, zero = kv_quant_params(x, L.q_bits, self.dtype)
                q = kv_quantize(x, scale, zero, L.q_bits, self.dtype)
                data = q if L.q_bits == 8 else pack_u4(q)
                p_base = (off + L.head_num * 2 * L.head_data_size() + head * 2 * L.head_param_size()
                          + is_v * L.head_param_size())
                pb = to_bits(np.stack([scale, zero], -1), self.dtype).view(np.uint8).reshape(S, 4)
                for s in range(S):
                    o = p_base + ti[s] * 4
                    self.pool[bi[s], o:o + 4] = pb[s]
            else:
                data = to_bits(x, self.dtype).view(np.uint8).reshape(S, -1)
            tds = L.token_data_size()
            for s in range(S):
                o = d_base + ti[s] * tds
                self.pool[bi[s], o:o + tds] = data[s]

    def read_tokens(self, block_ids, layer, head, n):
        """Gather + dequantise tokens 0..n-1 -> (K [n, D], V [n, D]) float32 exact in T."""
        L, off = self.layout, self.layer_offset(layer)
        D = L.head_dim
        pos = np.arange(n)
        bi = np.asarray(block_ids)[pos // L.block_len]
        ti = pos % L.block_len
        out = []
        tds = L.token_data_size()
        for is_v in (0, 1):
            d_base = off + head * 2 * L.head_data_size() + is_v * L.head_data_size()
            idx = d_base + ti[:, None] * tds + np.arange(tds)[None, :]
            raw = self.pool[bi[:, None], idx]                       # [n, tds] uint8
            if L.quant:
                p_base = (off + L.head_num * 2 * L.head_data_size() + head * 2 * L.head_param_size()
                          + is_v * L.head_param_size())
                pidx = p_base + ti[:, None] * 4 + np.arange(4)[None, :]
                pr = np.ascontiguousarray(self.pool[bi[:, None], pidx]).view(np.uint16)   # [n, 2]
                p = from_bits(pr, self.dtype)
                q = raw if L.q_bits == 8 else unpack_u4(raw, D)
                out.append(kv_dequant(q, p[:, 0], p[:, 1], self.dtype))
            else:
                out.append(from_bits(np.ascontiguousarray(raw).view(np.uint16), self.dtype))
        return out[0], out[1]


# --------------------------------------------------------------------------------------
# decode attention
# --------------------------------------------------------------------------------------
def decode_prologue(qkv, q_bias, k_bias, v_bias, history_len, num_heads, num_kv_heads, D, rope, dtype):
    """Bias + RoPE for the new token of each sequence (attention_universal.h:168-246).
    qkv: [B, (Hq + 2 Hkv) * D] float32 exact in T.  Returns q [B,Hq,D], k [B,Hkv,D], v [B,Hkv,D]."""
    B = qkv.shape[0]
    q = qkv[:, : num_heads * D].reshape(B, num_heads, D)
    k = qkv[:, num_heads * D: (num_heads + num_kv_heads) * D].reshape(B, num_kv_heads, D)
    v = qkv[:, (num_heads + num_kv_heads) * D:].reshape(B, num_kv_heads, D)
    if q_bias is not None:
        q = t_add(q, q_bias.reshape(1, num_heads, D), dtype)
        k = t_add(k, k_bias.reshape(1, num_kv_heads, D), dtype)
        v = t_add(v, v_bias.reshape(1, num_kv_heads, D), dtype)
    pos = np.asarray(history_len)[:, None]
    q = apply_rope(q, np.broadcast_to(pos, (B, num_heads)), rope, dtype)
    k = apply_rope(k, np.broadcast_to(pos, (B, num_kv_heads)), rope, dtype)
    return q, k, f32(v).copy()


def softmax_pv(q, K, V, inv_sqrt_dh, dtype):
    """One (sequence, q-head): q [D], K,V [n, D] -> out [D] in T, plus (M, L) for split tests.
    S = q.K^T (fp32 accumulate), P = exp2(S*c - M*c), L = sum(P) in fp32, P cast to T before
    PV (impl_81616.h:604-608), O fp32, out = (T)(O / L).   c = inv_sqrt_dh = log2(e)/sqrt(D)."""
    S = (K.astype(np.float64) @ q.astype(np.float64)).astype(np.float32)
    M = S.max()
    c = np.float32(inv_sqrt_dh)
    P = f32(np.exp2((S.astype(np.float64) - np.float64(M)) * np.float64(c)))
    L = P.astype(np.float64).sum()
    Pt = rn(P, dtype)
    O = Pt.astype(np.float64) @ V.astype(np.float64)
    return rn(f32(O / L), dtype), np.float32(M), np.float32(L), f32(O)


def decode_attention(qkv, cache: PagedKvCache, block_tables, history_len, layer, num_heads, num_kv_heads,
                     rope: RopeParam, inv_sqrt_dh=None, q_bias=None, k_bias=None, v_bias=None,
                     finished=None, readonly_block_num=None):
    """Full decode step for one layer: prologue, KV append (quantised) at position history_len,
    attention over history_len+1 tokens.  Returns out [B, Hq*D] float32 exact in T.
    Mutates `cache`.  (attention_universal.h:355-553, decoding_template.h:14-88)"""
    dtype = cache.dtype
    L = cache.layout
    D = L.head_dim
    B = qkv.shape[0]
    if inv_sqrt_dh is None:
        inv_sqrt_dh = np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D))))
    q, k, v = decode_prologue(qkv, q_bias, k_bias, v_bias, history_len, num_heads, num_kv_heads, D, rope, dtype)
    G = num_heads // num_kv_heads
    out = np.zeros((B, num_heads, D), np.float32)
    for b in range(B):
        if finished is not None and finished[b]:
            continue
        h_len = int(history_len[b])
        ro_len = int(readonly_block_num[b]) * L.block_len if readonly_block_num is not None else 0
        for h in range(num_kv_heads):
            if h_len >= ro_len:
                blk = block_tables[b][h_len // L.block_len]
                cache.write_token(blk, layer, h, h_len % L.block_len, k[b, h], v[b, h])
            K, V = cache.read_tokens(block_tables[b], layer, h, h_len + 1)
            for g in range(G):
                o, _, _, _ = softmax_pv(q[b, h * G + g], K, V, inv_sqrt_dh, dtype)
                out[b, h * G + g] = o
    return out.reshape(B, num_heads * D)


def merge_partials(Ms, Ls, Os, inv_sqrt_dh, dtype):
    """Split-K fix-up (reduce.cu:90-153):  O = sum_k w_k O_k / sum_k w_k L_k,
    w_k = exp2((M_k - M) * c).  Ms, Ls [k]; Os [k, D] (un-normalised fp32)."""
    M = np.max(Ms)
    w = f32(np.exp2((Ms.astype(np.float64) - np.float64(M)) * np.float64(np.float32(inv_sqrt_dh))))
    w = np.where(np.isneginf(Ms), np.float32(0), w)
    L = (w.astype(np.float64) * Ls.astype(np.float64)).sum()
    O = (w.astype(np.float64)[:, None] * Os.astype(np.float64)).sum(0)
    return rn(f32(O / L), dtype)


def process_kv(cache: PagedKvCache, block_table, layer, start, k, v, rope: RopeParam,
               k_bias=None, v_bias=None):
    """Prefill KV writer (kv_cache_utils_v2.cu:18-210 ProcessKV_v2): bias + RoPE on K, quantise, store
    tokens start..start+S-1.  k, v: [S, Hkv, D] float32 exact in T."""
    dtype = cache.dtype
    S, Hkv, D = k.shape
    if k_bias is not None:
        k = t_add(k, k_bias.reshape(1, Hkv, D), dtype)
        v = t_add(v, v_bias.reshape(1, Hkv, D), dtype)
    pos = (start + np.arange(S))[:, None]
    k = apply_rope(k, np.broadcast_to(pos, (S, Hkv)), rope, dtype)
    for h in range(Hkv):
        cache.write_tokens(block_table, layer, h, start, k[:, h], f32(v[:, h]))


# --------------------------------------------------------------------------------------
# prefill (context) attention
# --------------------------------------------------------------------------------------
def prefill_attention(qkv, cache: PagedKvCache, block_tables, q_lens, k_lens, layer, num_heads, num_kv_heads,
                      rope: RopeParam, inv_sqrt_dh=None, q_bias=None, k_bias=None, v_bias=None):
    """Chunked causal prefill for one layer: the reference writes the chunk's K/V into the pages
    (ProcessKV_v2, kv_cache_utils_v2.cu:18-210), reads them back de-quantised (FlattenKV_v2) and runs causal attention
    over keys [0, history + i] for query i of the chunk (attention_universal.h:355-553 with the prefill mainloop;
    same softmax arithmetic as decode).
    qkv: [sum(q_lens), (Hq + 2 Hkv) * D] float32 exact in T, sequences concatenated; history_b = k_lens[b] - q_lens[b].
    Returns out [sum(q_lens), Hq*D] float32 exact in T.  Mutates `cache`."""
    dtype = cache.dtype
    L = cache.layout
    D = L.head_dim
    if inv_sqrt_dh is None:
        inv_sqrt_dh = np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D))))
    G = num_heads // num_kv_heads
    T = qkv.shape[0]
    out = np.zeros((T, num_heads, D), np.float32)
    t0 = 0
    for b, (ql, kl) in enumerate(zip(q_lens, k_lens)):
        ql, kl = int(ql), int(kl)
        hist = kl - ql
        rows = qkv[t0:t0 + ql]
        q = rows[:, : num_heads * D].reshape(ql, num_heads, D)
        k = rows[:, num_heads * D: (num_heads + num_kv_heads) * D].reshape(ql, num_kv_heads, D)
        v = rows[:, (num_heads + num_kv_heads) * D:].reshape(ql, num_kv_heads, D)
        if q_bias is not None:
            q = t_add(q, q_bias.reshape(1, num_heads, D), dtype)
        pos = (hist + np.arange(ql))[:, None]
        q = apply_rope(q, np.broadcast_to(pos, (ql, num_heads)), rope, dtype)
        process_kv(cache, block_tables[b], layer, hist, k, v, rope, k_bias, v_bias)
        for h in range(num_kv_heads):
            K, V = cache.read_tokens(block_tables[b], layer, h, kl)
            for i in range(ql):
                n = hist + i + 1
                for g in range(G):
                    o, _, _, _ = softmax_pv(q[i, h * G + g], K[:n], V[:n], inv_sqrt_dh, dtype)
                    out[t0 + i, h * G + g] = o
        t0 += ql
    return out.reshape(T, num_heads * D)
