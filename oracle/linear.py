"""CPU oracle: weight-only-quant linear layers, RMSNorm, gated SiLU (test infrastructure only).

Restates
  lmdeploy/turbomind/weight_format.py:47-74,200-247   (AWQ unpack order, pack_u4_row)
  src/turbomind/kernels/gemm/cast.cu:134-165          (fuse_scales_and_zeros: (s, -z*s) in T)
  src/turbomind/kernels/gemm/transform.h:48-74        (w = __hfma((T)q, s, -z*s))
  src/turbomind/kernels/gemm/transform.h:92-97, models/linear_weight.cc:138-150 (fp8: w = bf16(e4m3) * bf16(scale))
  src/turbomind/kernels/gemm/epilogue.h:158-177       (gated SiLU: silu(x0) * x1 in fp32)
  src/turbomind/kernels/norm/rms_norm.cu:21-86,287-363, rms_norm_utils.cuh:6-15 (RMSNorm family)
  src/turbomind/kernels/quantization.cu:405-436       (QuantizeGroupwise: synthetic weight generator)
Accumulation is fp32 in the reference (mma f32 accumulate); summation order is not specified, so the
oracle accumulates in float64 and rounds once to fp32 -- GPU results are compared with a tolerance
stated in each test.
"""
import numpy as np

from .numerics import F16, BF16, rn, to_bits, from_bits, t_add, t_mul, t_fma, f32

AWQ_ORDER = [0, 4, 1, 5, 2, 6, 3, 7]


# ---- checkpoint-side packing (what an AWQ safetensors file holds) ----------------------------
def awq_pack(q):
    """q uint8 [K, N] (0..15) -> AWQ 'gemm' qweight int32 [K, N/8] (nibble i of a word = column
    8w + AWQ_ORDER^-1...).  Inverse of _unpack_awq_gemm (weight_format.py:56-60)."""
    K, N = q.shape
    g = q.reshape(K, N // 8, 8).astype(np.uint32)
    w = np.zeros((K, N // 8), np.uint32)
    # _unpack_awq_gemm: ys[j] = xs[order[j]]  => column j of the group comes from nibble order[j]
    for j, nib in enumerate(AWQ_ORDER):
        w |= g[:, :, j] << np.uint32(4 * nib)
    return w.view(np.int32)


def awq_unpack(w):
    """_unpack_awq_gemm (weight_format.py:56-60): int32 [K, N/8] -> uint8 [K, N]."""
    w = np.ascontiguousarray(w).view(np.uint32)
    K, N8 = w.shape
    out = np.zeros((K, N8, 8), np.uint8)
    for j, nib in enumerate(AWQ_ORDER):
        out[:, :, j] = (w >> np.uint32(4 * nib)) & 0xF
    return out.reshape(K, N8 * 8)


def pack_u4_row(q):
    """pack_u4_row (weight_format.py:63-74): uint8 [K, N] -> int32 [K, N/8], nibble i = column 8w+i.
    This is the loader-visible 'pre-prepare' TurboMind layout handed to LinearWeight."""
    K, N = q.shape
    g = q.reshape(K, N // 8, 8).astype(np.uint32)
    w = np.zeros((K, N // 8), np.uint32)
    for i in range(8):
        w |= g[:, :, i] << np.uint32(4 * i)
    return w.view(np.int32)


def unpack_u4_row(w):
    w = np.ascontiguousarray(w).view(np.uint32)
    K, N8 = w.shape
    out = np.zeros((K, N8, 8), np.uint8)
    for i in range(8):
        out[:, :, i] = (w >> np.uint32(4 * i)) & 0xF
    return out.reshape(K, N8 * 8)


# ---- synthetic weight generator (QuantizeGroupwise, quantization.cu:405-436) ------------------
def quantize_groupwise_u4(w, group=128):
    """w float32 [K, N] -> (q uint8 [K,N], scales f16 [K/g, N], zeros f16 [K/g, N]) asymmetric u4."""
    K, N = w.shape
    wg = w.reshape(K // group, group, N).astype(np.float32)
    mn = wg.min(1)
    mx = wg.max(1)
    scale = np.maximum(mx - mn, np.float32(1e-5)) / np.float32(15)
    zero = np.clip(-np.rint(mn / scale), 0, 15)
    q = np.clip(np.rint(wg / scale[:, None]) + zero[:, None], 0, 15).astype(np.uint8).reshape(K, N)
    return q, rn(scale, F16), rn(zero, F16)


# ---- dequant ---------------------------------------------------------------------------------
def dequant_u4(q, scales, zeros, dtype=F16):
    """w[k,n] = __hfma((T)q, s, t),  t = (T)(-z * s)   (cast.cu:155, transform.h:48-74)."""
    K, N = q.shape
    g = K // scales.shape[0]
    s = np.repeat(f32(scales), g, axis=0)
    t = np.repeat(rn(-f32(zeros) * f32(scales), dtype), g, axis=0)
    return t_fma(q.astype(np.float32), s, t, dtype)


_E4M3_LUT = None


def e4m3_to_f32(b):
    """float8_e4m3fn bits -> float32."""
    global _E4M3_LUT
    if _E4M3_LUT is None:
        lut = np.zeros(256, np.float32)
        for i in range(256):
            s = -1.0 if i & 0x80 else 1.0
            e = (i >> 3) & 0xF
            m = i & 7
            if e == 0xF and m == 7:
                v = np.nan
            elif e == 0:
                v = s * (m / 8.0) * 2.0 ** (-6)
            else:
                v = s * (1 + m / 8.0) * 2.0 ** (e - 7)
            lut[i] = v
        _E4M3_LUT = lut
    return _E4M3_LUT[np.asarray(b, np.uint8)]


def f32_to_e4m3(x):
    """RNE float32 -> e4m3fn bits (saturating to +-448), for the synthetic generator."""
    lut = e4m3_to_f32(np.arange(256, dtype=np.uint8))
    pos = lut[:127].astype(np.float64)              # 0 .. 448 ascending
    ax = np.minimum(np.abs(np.asarray(x, np.float64)), 448.0)
    idx = np.searchsorted(pos, ax)
    idx = np.clip(idx, 1, 126)
    lo, hi = pos[idx - 1], pos[idx]
    pick_hi = (ax - lo > hi - ax) | ((ax - lo == hi - ax) & ((idx & 1) == 0))
    r = np.where(pick_hi, idx, idx - 1).astype(np.uint8)
    return np.where(np.signbit(x), r | 0x80, r).astype(np.uint8)


def dequant_fp8(w8, scale_inv, block=128, dtype=BF16):
    """Non-sm90 fp8-weight path: w = bf16(e4m3) *_bf16 bf16(scale_block); 128x128 block scales
    (f32 in the checkpoint) expanded to per-K-group scales in T (linear_weight.cc:138-150)."""
    K, N = w8.shape
    s = rn(f32(scale_inv), dtype)                                   # [K/128, N/128] cast to T
    s = np.repeat(np.repeat(s, block, axis=0), block, axis=1)[:K, :N]
    return t_mul(rn(e4m3_to_f32(w8), dtype), s, dtype)


def dequant_fp8_cols(w8, scales, group=128, dtype=BF16):
    """Same arithmetic with the block scales already expanded along N (what the reference's kernel consumes after
    linear_weight.cc:138-150): scales f32 [K/group, N]."""
    K, N = w8.shape
    s = np.repeat(rn(f32(scales), dtype), group, axis=0)[:K]
    return t_mul(rn(e4m3_to_f32(w8), dtype), s, dtype)


# ---- linear / epilogues ----------------------------------------------------------------------
def linear(x, w, dtype, bias=None):
    """D[M,N] = (T)(sum_k x[m,k] w[k,n]) fp32-accumulated (oracle: float64 accumulate)."""
    acc = x.astype(np.float64) @ w.astype(np.float64)
    if bias is not None:
        acc = acc + bias.astype(np.float64)
    return rn(acc.astype(np.float32), dtype)


def linear_gated_silu(x, w_interleaved, dtype):
    """Fused w1w3 GEMM + GatedActivation<Silu> (epilogue.h:158-177): columns are element-interleaved
    (gate, up); out[m, j] = (T)(silu(acc[m,2j]) * acc[m,2j+1]) in fp32, silu(x) = x / (1 + exp(-x))."""
    acc = (x.astype(np.float64) @ w_interleaved.astype(np.float64)).astype(np.float32)
    g, u = acc[:, 0::2], acc[:, 1::2]
    with np.errstate(over="ignore"):
        silu = f32(g / f32(np.float32(1) + f32(np.exp(-g.astype(np.float64)))))
    return rn(f32(silu * u), dtype)


def silu_mul(gate, up, dtype):
    """Unfused activation (kernels/activation.cu): computed in fp32 from T inputs."""
    g, u = f32(gate), f32(up)
    with np.errstate(over="ignore"):
        silu = f32(g / f32(np.float32(1) + f32(np.exp(-g.astype(np.float64)))))
    return rn(f32(silu * u), dtype)


def rmsnorm(x, w, eps, dtype):
    """RMSNorm (rms_norm.cu:21-86): sum x^2 in fp32, r = rsqrt(mean + eps), y = (T)((float)x * r) * w."""
    x = f32(x)
    ss = (x.astype(np.float64) ** 2).sum(-1, keepdims=True)
    r = f32(1.0 / np.sqrt(f32(f32(ss / x.shape[-1]) + np.float32(eps)).astype(np.float64)))
    return t_mul(rn(f32(x * r), dtype), f32(w), dtype)


def qk_rmsnorm(qkv, q_w, k_w, num_heads, num_kv_heads, head_dim, eps, dtype):
    """Per-head RMSNorm of the q and k heads of fused QKV rows (Qwen3 family), as the reference's release build computes it
    (invokeQkRMSNorm, rms_norm.cu:139-260; instruction order read from the compiled kernel): 16 lanes x 8 elements per
    head; per lane an fp32 FMA chain over its 8 elements starting from 0; butterfly over lane distances 8, 4, 2, 1;
    r = rsqrt(fma(sum, 1/dim, eps)); y = T(float(x) * r) *_T w.  (The hardware's rsqrt is an approximation: r may differ
    from this restatement in its last fp32 bit, which moves an output by at most one T ulp.)  v is left untouched."""
    assert head_dim == 128
    qkv = f32(qkv).copy()
    T = qkv.shape[0]
    nqk = (num_heads + num_kv_heads) * head_dim
    x = qkv[:, :nqk].reshape(T, num_heads + num_kv_heads, 16, 8)
    s = np.zeros(x.shape[:3], np.float32)
    for i in range(8):                                   # fma(x, x, s): one rounding (the product is exact in float64)
        s = (x[..., i].astype(np.float64) ** 2 + s.astype(np.float64)).astype(np.float32)
    for d in (8, 4, 2, 1):                               # butterfly: every lane ends with the same sum
        s = f32(s + s[..., np.arange(16) ^ d])
    ss = s[..., :1]
    arg = (ss.astype(np.float64) * float(np.float32(1.0 / head_dim)) + float(np.float32(eps))).astype(np.float32)
    r = f32(1.0 / np.sqrt(arg.astype(np.float64)))[..., None]
    w = np.concatenate([np.broadcast_to(f32(q_w), (num_heads, head_dim)),
                        np.broadcast_to(f32(k_w), (num_kv_heads, head_dim))]).reshape(1, -1, 16, 8)
    y = t_mul(rn(f32(x * r), dtype), w, dtype)
    qkv[:, :nqk] = y.reshape(T, nqk)
    return qkv


def residual_rmsnorm(hidden, residual, w, eps, dtype, bias=None):
    """BiasResidualRMSNorm (rms_norm.cu:287-363): residual <- residual + hidden (+bias) in T;
    hidden <- RMSNorm(residual) * w.  Returns (hidden_out, residual_out)."""
    r = t_add(residual, hidden, dtype)
    if bias is not None:
        r = t_add(r, bias, dtype)
    return rmsnorm(r, w, eps, dtype), r


def allreduce_residual_rmsnorm(partials, residual, w, eps, dtype, bias=None):
    """TP fused op (comm/device_comm.h:88-101): hidden = sum_ranks partials (rank order, T adds as in
    the push/pull kernels, fused_allreduce.cu:25-167), then residual_rmsnorm."""
    acc = f32(partials[0])
    for p in partials[1:]:
        acc = t_add(acc, p, dtype)
    return residual_rmsnorm(acc, residual, w, eps, dtype, bias)
