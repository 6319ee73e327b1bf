"""The torch-CPU baseline port (oracle/cpu_baseline.py, bench.py's `cpu_baseline` / `--impl reference` leg) against the
numpy oracle: the baseline computes in fp32, the oracle in the reference's T arithmetic, so every building block must
agree to fp16 rounding.  CPU only; tiny shapes."""
import math
from types import SimpleNamespace

import numpy as np
import torch

from oracle import attention as OA
from oracle import cpu_baseline as CB
from oracle import linear as OL
from oracle.numerics import F16


def test_unpack_matches_oracle_row_packing():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, (256, 64), dtype=np.uint8)
    packed = OL.pack_u4_row(q)
    got = CB.unpack_u4_rows(torch.from_numpy(packed)).numpy()
    assert np.array_equal(got, q)


def test_awq_linear_matches_oracle_dequant_matmul():
    rng = np.random.default_rng(1)
    K, N, M = 256, 128, 5
    w = (rng.standard_normal((K, N)) * 0.05).astype(np.float32)
    q, s, z = OL.quantize_groupwise_u4(w)
    x = (rng.standard_normal((M, K))).astype(np.float16).astype(np.float32)
    ref = OL.linear(x, OL.dequant_u4(q, s, z, F16), F16)
    got = CB.awq_linear(torch.from_numpy(x), torch.from_numpy(OL.pack_u4_row(q)), torch.from_numpy(s.astype(np.float32)),
                        torch.from_numpy(z.astype(np.float32))).numpy()
    # fp32 (q - z) * s vs fp16 fma(q, s, -z*s): one fp16 rounding per weight, sqrt(K)-averaged in the dot product
    tol = 2.0 ** -10 * np.abs(x).max() * np.abs(w).max() * math.sqrt(K) * 4 + 2.0 ** -10 * np.abs(ref).max()
    assert np.abs(got - ref).max() <= tol


def test_rmsnorm_matches_oracle():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 512)).astype(np.float16).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(512)).astype(np.float16).astype(np.float32)
    ref = OL.rmsnorm(x, w, 1e-5, F16)
    got = CB.rmsnorm(torch.from_numpy(x), torch.from_numpy(w), 1e-5).numpy()
    assert np.abs(got - ref).max() <= 2.0 ** -9 * np.abs(ref).max()      # two fp16 roundings in the oracle


def test_rope_matches_oracle():
    rng = np.random.default_rng(3)
    B, H, D = 4, 3, 128
    x = rng.standard_normal((B, H, D)).astype(np.float16).astype(np.float32)
    pos = np.array([0, 1, 777, 4095])
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=D, base=10000.0)
    ref = OA.apply_rope(x, np.broadcast_to(pos[:, None], (B, H)), rope, F16)
    got = CB.rope_adjacent(torch.from_numpy(x), torch.from_numpy(pos), 10000.0, D).numpy()
    # cos/sin rounded to fp16 in the oracle (2^-11 relative) + product roundings; |x| <= ~4.5
    assert np.abs(got - ref).max() <= 3 * 2.0 ** -10 * np.abs(x).max()


def test_layer_attention_block_matches_plain_softmax():
    """The gather/dequant/softmax inside CpuLayer.forward is checked against a direct per-(batch, head) evaluation on
    the un-permuted cache: the permutation (the paged gather model) must not change the result."""
    gen = torch.Generator().manual_seed(5)
    hidden, heads, kv_heads, inter, D, B, ctx = 256, 2, 1, 256, 128, 2, 48
    layer = CB.CpuLayer(hidden, heads, kv_heads, inter, "fp16", gen, D)
    kv = dict(kq=torch.randint(0, 256, (B, kv_heads, ctx, D), generator=gen, dtype=torch.uint8),
              vq=torch.randint(0, 256, (B, kv_heads, ctx, D), generator=gen, dtype=torch.uint8),
              ks=torch.rand(B, kv_heads, ctx, generator=gen) * 0.03, vs=torch.rand(B, kv_heads, ctx, generator=gen) * 0.03,
              kz=-torch.rand(B, kv_heads, ctx, generator=gen) * 4, vz=-torch.rand(B, kv_heads, ctx, generator=gen) * 4)
    residual = torch.randn(B, hidden, generator=gen)
    hist = torch.full((B,), ctx, dtype=torch.int64)
    outs = []
    for perm in (torch.arange(ctx), torch.randperm(ctx, generator=gen)):
        kvp = dict(kv, perm=perm)
        # softmax over the token axis is permutation invariant when K and V are gathered with the same indices
        outs.append(layer.forward(residual, kvp, hist, 10000.0))
    assert torch.allclose(outs[0], outs[1], rtol=1e-4, atol=1e-5)
    assert torch.isfinite(outs[0]).all()


def test_time_decode_reports_full_model_extrapolation():
    spec = SimpleNamespace(hidden=256, heads=2, kv_heads=1, inter=256, layers=3, vocab=512, rope_theta=10000.0,
                           weight_format="awq")
    r = CB.time_decode(spec, batch=2, ctx=64, layers_sample=1, reps=1, threads=1)
    assert r["threads"] == 1 and r["t_layer"] > 0 and r["t_head"] > 0
    assert math.isclose(r["t_step"], r["t_layer"] * 3 + r["t_head"], rel_tol=1e-9)
    assert math.isclose(r["tokens_per_s"], 2 / r["t_step"], rel_tol=1e-9)
    assert "1 of 3 decoder layers" in r["sample"]


def test_usable_cores_is_bounded():
    n = CB.usable_cores()
    assert 1 <= n <= 32
    assert CB.usable_cores(cap=1) == 1
