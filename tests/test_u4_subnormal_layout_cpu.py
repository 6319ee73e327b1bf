"""The f16 "subnormal" u4 weight layout (lmdeploy_b200/csrc/gemm.cuh: kU4FastShift, gemm.cu: u4_fast_*_kernel) restated in numpy
and checked exhaustively on the CPU: the three facts that make  fma(q', s', t) == fma(q, s, t)  bit for bit.

  (a) a nibble q placed at bits 7..10 of an f16 lane is the f16 number q * 2^-17 (the encoding is linear over the subnormals and
      the first binade), for every q;
  (b) every f16 scale with |s| * 2^17 <= 65504 (what b200_linear_prepare checks) has s * 2^17 exactly representable in f16,
      so the product q' * s' is the same real number as q * s and the single rounding of the fma sees the same value;
  (c) the packed words stored rotated left by 3 yield, through the kernel's three shifts / one rotate and the mask 0x07800780,
      the same (n0,n4) (n1,n5) (n2,n6) (n3,n7) pairs the plain layout's lop3 extraction yields.

The device side of the same statement is tests/test_linear_gpu.py (both layouts against the oracle) and the parity tests
against the reference's own GEMM."""
import numpy as np

SHIFT = 17
MASK = np.uint32(0x07800780)


def test_nibble_at_bits_7_to_10_is_q_times_2_pow_minus_17():
    q = np.arange(16, dtype=np.uint16)
    h = (q << 7).view(np.float16)
    assert (h.astype(np.float64) == q.astype(np.float64) * 2.0 ** -SHIFT).all()
    # and it stops being linear one bit higher (why bits 7..10 and not 8..11): bit 11 is the second exponent bit
    assert ((q << 8).view(np.float16).astype(np.float64) != q.astype(np.float64) * 2.0 ** -16).any()


def test_every_admitted_scale_times_2_pow_17_is_exact_and_the_fma_sees_the_same_real_number():
    bits = np.arange(1 << 16, dtype=np.uint16)
    s = bits.view(np.float16)
    s64 = s.astype(np.float64)
    with np.errstate(invalid="ignore"):
        ok = np.isfinite(s64) & (np.abs(s64) * 2.0 ** SHIFT <= 65504.0)      # the check kernel's predicate
    assert ok.sum() > 28000 and np.abs(s64[ok]).max() < 0.5                  # every scale below 0.5 (subnormals too) is admitted
    sp = (s64[ok] * 2.0 ** SHIFT).astype(np.float16)                         # what prepare stores (round to nearest)
    assert (sp.astype(np.float64) == s64[ok] * 2.0 ** SHIFT).all()          # ... exactly
    # fma(q', s', t) vs fma(q, s, t) with t = RN(-z * s): float64 holds q*s + t exactly (11 + 4 bit factors, f16 exponent
    # range), so one conversion to f16 IS the correctly rounded fma
    sel = np.nonzero(ok)[0][:: 7]
    s_ = s64[sel][:, None, None]
    sp_ = (s_ * 2.0 ** SHIFT).astype(np.float16).astype(np.float64)
    q = np.arange(16, dtype=np.float64)[None, :, None]
    z = np.arange(16, dtype=np.float64)[None, None, :]
    t = (-z * s_).astype(np.float16).astype(np.float64)
    plain = (q * s_ + t).astype(np.float16)
    fast = ((q * 2.0 ** -SHIFT) * sp_ + t).astype(np.float16)
    assert (plain.view(np.uint16) == fast.view(np.uint16)).all()


def _rotl(x, n):
    x = x.astype(np.uint64)
    return (((x << np.uint64(n)) | (x >> np.uint64(32 - n))) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def test_rotated_words_unpack_to_the_same_nibble_pairs():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 1 << 32, 200000, dtype=np.uint64).astype(np.uint32)
    x[:4] = [0, 0xFFFFFFFF, 0x80000001, 0x0F0F0F0F]
    n = [(x >> np.uint32(4 * i)) & np.uint32(15) for i in range(8)]
    xr = _rotl(x, 3)                                                    # u4_fast_apply_kernel
    h = [(xr << np.uint32(4)) & MASK, xr & MASK, (xr >> np.uint32(4)) & MASK, _rotl(xr, 24) & MASK]   # rotr 8 == rotl 24
    for j in range(4):
        lo = (h[j] & np.uint32(0xFFFF)) >> np.uint32(7)
        hi = (h[j] >> np.uint32(16)) >> np.uint32(7)
        assert (lo == n[j]).all() and (hi == n[j + 4]).all()           # (n_j, n_{j+4}) like TypeOps<half>::cvt_u4x8
