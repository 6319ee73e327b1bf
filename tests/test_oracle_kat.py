"""Known-answer tests that pin the CPU oracle to the reference's bit-level definitions.

The reference ships no golden vectors for this path; these KATs are derived by
hand from the PTX-level definitions in the reference sources (cited per test)."""
import numpy as np

from oracle import attention as OA
from oracle import linear as OL
from oracle.numerics import F16, BF16, rn, to_bits, from_bits, t_fma


def test_layout_sizes_match_survey():
    # block.h:172-176 ; SURVEY 8a2: 7B kv-int8: 8*2*64*(128+4) = 135168 ; llama3-8B kv-int4: 69632 ; fp16: 262144
    assert OA.KvLayout(8, 64, 128, 8).layer_size() == 135168
    assert OA.KvLayout(8, 64, 128, 4).layer_size() == 69632
    assert OA.KvLayout(8, 64, 128, 16).layer_size() == 262144
    L = OA.KvLayout(8, 64, 128, 8)
    # data region [head][K|V][token][D], params after all data (block.h:178-198)
    assert L.k_data(1, 0) == 2 * 64 * 128 and L.v_data(0, 0) == 64 * 128
    assert L.k_param(0, 0) == 8 * 2 * 64 * 128 and L.v_param(0, 3) == L.k_param(0, 3) + 64 * 4


def test_u4_nibble_order():
    # ConvertKvCache<T,uint4_t>::pack (quantization.h:461-474): nibble j holds element [0,2,4,6,1,3,5,7][j]
    q = np.arange(8, dtype=np.uint8)[None]
    b = OA.pack_u4(q)
    w = int(b.view(np.uint32)[0, 0])
    assert [(w >> (4 * j)) & 15 for j in range(8)] == [0, 2, 4, 6, 1, 3, 5, 7]
    assert (OA.unpack_u4(b, 8) == q).all()
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, (5, 128)).astype(np.uint8)
    assert (OA.unpack_u4(OA.pack_u4(q), 128) == q).all()


def test_kv_quant_kat_u8():
    # x = 0..127 * 2: min 0, max 254 -> scale = 254 * f32(1/255) -> f16 ; inv_scale = f16(1/scale)
    x = (np.arange(128, dtype=np.float32) * 2)[None]
    s, z = OA.kv_quant_params(x, 8, F16)
    assert z[0] == 0.0
    assert s[0] == np.float32(np.float16(np.float32(254.0) * (np.float32(1) / np.float32(255))))
    q = OA.kv_quantize(x, s, z, 8, F16)
    d = OA.kv_dequant(q, s, z, F16)
    assert np.abs(d - x).max() <= float(s[0])          # within one quantisation step
    # saturation + RNE: (x - zero) * inv_scale evaluated in f16, cvt.rni.sat.u8
    inv = np.float16(np.float32(1) / np.float32(s[0]))
    ref = np.clip(np.rint((x.astype(np.float16) * inv).astype(np.float32)), 0, 255).astype(np.uint8)
    assert (q == ref).all()


def test_kv_quant_constant_vector_edge_case():
    # constant head vector -> scale 0 -> inv_scale inf -> (x-zero)*inf = NaN -> 0
    x = np.full((1, 128), 1.5, np.float32)
    s, z = OA.kv_quant_params(x, 8, F16)
    assert s[0] == 0 and z[0] == 1.5
    q = OA.kv_quantize(x, s, z, 8, F16)
    assert (q == 0).all()
    assert (OA.kv_dequant(q, s, z, F16) == 1.5).all()


def test_kv_quant_u4_clamp():
    x = np.linspace(-1, 1, 128, dtype=np.float32)[None]
    x = rn(x, F16)
    s, z = OA.kv_quant_params(x, 4, F16)
    q = OA.kv_quantize(x, s, z, 4, F16)
    assert q.max() <= 15 and q.min() == 0


def test_hfma_single_rounding():
    # __hfma: one rounding.  255 * 0.0123 - 1.5 evaluated exactly then rounded to f16
    a, b, c = np.float32(255), rn(np.float32(0.0123), F16), np.float32(-1.5)
    exact = float(a) * float(b) + float(c)
    assert t_fma(a, b, c, F16) == np.float32(np.float16(exact))


def test_bf16_rounding():
    # RNE ties: 1 + 2^-8 is a tie between 1.0 and 1+2^-7 -> even (1.0); 1 + 3*2^-8 -> 1 + 2^-6
    assert rn(np.float32(1 + 2 ** -8), BF16) == 1.0
    assert rn(np.float32(1 + 3 * 2 ** -8), BF16) == np.float32(1 + 2 ** -6)
    x = np.float32(3.14159)
    assert from_bits(to_bits(rn(x, BF16), BF16), BF16) == rn(x, BF16)


def test_rope_kernel_params():
    # init_rope_kernel_param (attention_weight.cc:37-93)
    r = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e6)
    kp = r.kernel_params()
    assert kp["scale_factor"] == np.float32(-np.log2(np.float32(1e6)) / np.float32(128))
    assert kp["inv_factor"] == 1.0
    f = OA.rope_inv_freq(r)
    assert f[0] == 1.0 and abs(f[-1] - 1e6 ** (-126 / 128)) / f[-1] < 1e-6
    r3 = OA.RopeParam(type=OA.ROPE_LLAMA3, dim=128, base=5e5, factor=8.0)
    f3 = OA.rope_inv_freq(r3)
    f0 = OA.rope_inv_freq(OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=5e5))
    # high-frequency dims unchanged, low-frequency dims divided by factor
    assert f3[0] == f0[0] and abs(f3[-1] - f0[-1] / 8) / f3[-1] < 1e-6


def test_rope_matches_hf_half_rotation_after_reorder():
    """TurboMind rotates adjacent pairs of *reordered* channels; HF rotates (i, i+D/2) pairs
    (lmdeploy/pytorch/backends/default/apply_rotary_emb.py:8-76; models/utils.py:306-324 reorder)."""
    rng = np.random.default_rng(1)
    D = 128
    x = rn(rng.standard_normal(D).astype(np.float32), F16)
    r = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=D, base=1e4)
    pos = 37
    # reorder_rotary_emb: interleave the two halves
    xi = np.empty(D, np.float32)
    xi[0::2], xi[1::2] = x[: D // 2], x[D // 2:]
    y = OA.apply_rope(xi[None], np.array([pos]), r, F16)[0]
    inv = 1.0 / (1e4 ** (np.arange(0, D, 2) / D))
    c, s = np.cos(pos * inv), np.sin(pos * inv)
    hf = np.concatenate([x[: D // 2] * c - x[D // 2:] * s, x[D // 2:] * c + x[: D // 2] * s])
    got = np.concatenate([y[0::2], y[1::2]])
    assert np.abs(got - hf).max() < 4e-3


def test_awq_pack_roundtrip_and_order():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, (16, 64)).astype(np.uint8)
    w = OL.awq_pack(q)
    assert (OL.awq_unpack(w) == q).all()
    # AWQ order [0,4,1,5,2,6,3,7]: column j of a group sits in nibble order[j] (weight_format.py:56-60)
    one = np.zeros((1, 8), np.uint8)
    one[0, 1] = 0xF
    assert int(OL.awq_pack(one).view(np.uint32)[0, 0]) == 0xF << (4 * 4)
    assert (OL.unpack_u4_row(OL.pack_u4_row(q)) == q).all()
    assert int(OL.pack_u4_row(one).view(np.uint32)[0, 0]) == 0xF << 4


def test_w4_dequant_matches_pytorch_engine_semantics():
    """lmdeploy/pytorch/backends/default/awq_modules.py:11-46: w = (q - z) * s ; TurboMind computes
    fma(q, s, -(z*s)) in f16: identical up to f16 rounding of the two forms."""
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((256, 64)) * 0.02).astype(np.float32)
    q, s, z = OL.quantize_groupwise_u4(w, 128)
    d = OL.dequant_u4(q, s, z, F16)
    ref = (q.astype(np.float32) - np.repeat(z, 128, 0)) * np.repeat(s, 128, 0)
    assert np.abs(d - ref).max() <= np.abs(ref).max() * 2 ** -9
    assert np.abs(d - w).max() <= float(s.max()) * 0.51 + 1e-4


def test_e4m3_table():
    assert OL.e4m3_to_f32(np.uint8(0x7E)) == 448.0 and OL.e4m3_to_f32(np.uint8(0x08)) == 2.0 ** -6
    assert OL.e4m3_to_f32(np.uint8(0x01)) == 2.0 ** -9 and np.isnan(OL.e4m3_to_f32(np.uint8(0x7F)))
    x = np.array([0.0, 1.0, -1.0, 448.0, 1000.0, 0.0175, -3.3], np.float32)
    b = OL.f32_to_e4m3(x)
    back = OL.e4m3_to_f32(b)
    assert back[0] == 0 and back[1] == 1 and back[2] == -1 and back[3] == 448 and back[4] == 448
    assert abs(back[6] + 3.25) < 1e-6


def test_softmax_split_merge_consistency():
    """Split-K merge (reduce.cu:90-153) of per-split (M, L, O) equals the unsplit result."""
    rng = np.random.default_rng(3)
    n, D = 300, 128
    q = rn(rng.standard_normal(D).astype(np.float32), F16)
    K = rn(rng.standard_normal((n, D)).astype(np.float32), F16)
    V = rn(rng.standard_normal((n, D)).astype(np.float32), F16)
    c = np.float32(1.4426950408889634 / np.sqrt(128))
    full, _, _, _ = OA.softmax_pv(q, K, V, c, F16)
    Ms, Ls, Os = [], [], []
    for a, b in ((0, 64), (64, 200), (200, 300)):
        _, M, L, O = OA.softmax_pv(q, K[a:b], V[a:b], c, F16)
        Ms.append(M), Ls.append(L), Os.append(O)
    merged = OA.merge_partials(np.array(Ms), np.array(Ls), np.stack(Os), c, F16)
    assert np.abs(merged - full).max() <= 2e-3


def test_rmsnorm_and_gated_silu():
    rng = np.random.default_rng(0)
    x = rn(rng.standard_normal((3, 256)).astype(np.float32), F16)
    w = rn((1 + 0.02 * rng.standard_normal(256)).astype(np.float32), F16)
    y = OL.rmsnorm(x, w, 1e-5, F16)
    ref = x / np.sqrt((x.astype(np.float64) ** 2).mean(-1, keepdims=True) + 1e-5) * w
    assert np.abs(y - ref).max() < 5e-3
    h, r = OL.residual_rmsnorm(x, x, w, 1e-5, F16)
    assert (r == rn(x + x, F16)).all()
    g = OL.silu_mul(x, x, F16)
    assert np.abs(g - x * x / (1 + np.exp(-x))).max() < 1e-2
