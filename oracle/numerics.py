"""Rounding helpers: fp16 / bf16 arithmetic emulated on float32/float64 numpy arrays.

Values of type T (half or bfloat16) are carried as float32 arrays that are exactly
representable in T.  Every T-typed operation of the reference (operator+,-,* on
half/nv_bfloat16, __hfma) is a single correctly-rounded (RNE) operation; we compute
in a wider type where the result is exact (or innocuously double-rounded: p_wide >=
2*p_T+2) and round once.
"""
import numpy as np

F16 = "f16"
BF16 = "bf16"


def rn_bf16(x):
    """Round float32/float64 array to nearest-even bfloat16, returned as float32."""
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32)
    nan = np.isnan(x)
    if np.any(nan):
        out = np.where(nan, np.float32(np.nan), out)
    return out.reshape(x.shape)


def rn(x, dtype):
    """Round to T (RNE) and return as float32."""
    if dtype == F16:
        with np.errstate(over="ignore", invalid="ignore"):
            return np.asarray(x).astype(np.float16).astype(np.float32)
    if dtype == BF16:
        x = np.asarray(x)
        if x.dtype == np.float64:
            # double -> bf16 directly (avoid double rounding through f32):
            # round-to-odd into f32 then RNE to bf16
            f = x.astype(np.float32)
            err = f.astype(np.float64) - x
            u = f.view(np.uint32)
            # if inexact and f32 mantissa even, nudge toward the true value's side to make it odd
            inexact = err != 0
            even = (u & 1) == 0
            toward = np.where((err > 0) == (f > 0), -1, 1).astype(np.int64)  # move magnitude toward x
            u2 = np.where(inexact & even, (u.astype(np.int64) + toward).astype(np.uint32), u)
            f = u2.view(np.float32)
            return rn_bf16(f)
        return rn_bf16(x)
    raise ValueError(dtype)


def to_bits(x, dtype):
    """float32 array (exact in T) -> uint16 storage bits."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == F16:
        return x.astype(np.float16).view(np.uint16)
    return (x.view(np.uint32) >> 16).astype(np.uint16)


def from_bits(u, dtype):
    u = np.asarray(u, dtype=np.uint16)
    if dtype == F16:
        return u.view(np.float16).astype(np.float32)
    return (u.astype(np.uint32) << 16).view(np.float32)


def t_add(a, b, dtype):
    return rn(np.asarray(a, np.float32) + np.asarray(b, np.float32), dtype)


def t_sub(a, b, dtype):
    return rn(np.asarray(a, np.float32) - np.asarray(b, np.float32), dtype)


def t_mul(a, b, dtype):
    return rn(np.asarray(a, np.float32) * np.asarray(b, np.float32), dtype)


def t_fma(a, b, c, dtype):
    """__hfma(a,b,c): single rounding.  a*b is exact in float64 (<=22 significant bits)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    c = np.asarray(c, np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        return rn(a * b + c, dtype)


def f32(x):
    return np.asarray(x, dtype=np.float32)
