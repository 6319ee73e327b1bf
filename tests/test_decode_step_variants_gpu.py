"""Whole decode steps beyond the C2 default: bf16 (dense / fp8 weights), synthetic fp8 model, q/k/v bias (Qwen2), odd GQA
groups, per-head q/k RMSNorm (Qwen3) -- each against the oracle.  First device run: round 2."""
import numpy as np
import pytest

from oracle import linear as OL
from oracle.numerics import BF16, F16, rn

from test_linear_gpu import DevLinear, _assert_close

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------------
# whole decode step in bf16: fp8 weights (Llama-3-70B FP8 is BASELINE config C5) and dense bf16 weights
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kv_bits", [8, 16])
@pytest.mark.parametrize("fmt", ["fp8", "fp16"])
def test_decode_step_bf16_matches_oracle(kv_bits, fmt):
    """Same protocol as test_decode_step_gpu.test_decode_step_matches_oracle with T = bf16.  Bars: bf16 keeps 8
    significand bits (8x the fp16 rounding), so logits |diff| <= 3.2e-2 and token ids must agree wherever the oracle's
    top-1/top-2 margin exceeds twice that."""
    import torch
    from oracle import attention as OA
    from oracle.decode_step import TinyWeights, decode_step
    from harness import from_dev_T
    from test_decode_step_gpu import build_device_model
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=1024, vocab=1024, fmt=fmt, dtype=BF16, seed=kv_bits)
    hist = np.array([0, 5, 64, 130, 311], np.int32)
    B = len(hist)
    rng = np.random.default_rng(1)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    layout = OA.KvLayout(2, 64, 128, kv_bits)
    blocks_per = [(int(h) + 1 + 63) // 64 for h in hist]
    nblk = sum(blocks_per) + 2
    cache = OA.PagedKvCache(nblk, w.layers, layout, BF16)
    perm = rng.permutation(nblk)
    tables, o = [], 0
    for n in blocks_per:
        tables.append(perm[o:o + n].tolist())
        o += n
    for l in range(w.layers):
        for b in range(B):
            h = int(hist[b])
            if h:
                k = OA.rn(rng.standard_normal((h, 2, 128)).astype(np.float32), BF16)
                v = OA.rn(rng.standard_normal((h, 2, 128)).astype(np.float32), BF16)
                OA.process_kv(cache, tables[b], l, 0, k, v, rope)
    tokens = rng.integers(0, w.vocab, B).astype(np.int32)

    m = build_device_model(w, 0 if kv_bits == 16 else kv_bits, max_batch=8)
    pool = torch.from_numpy(cache.pool.copy()).cuda()
    ptrs = torch.tensor([pool.data_ptr() + int(b) * cache.block_bytes for t in tables for b in t], dtype=torch.int64,
                        device="cuda")
    cu_blocks = torch.tensor(np.concatenate([[0], np.cumsum(blocks_per)]), dtype=torch.int32, device="cuda")
    cu_q = torch.arange(B + 1, dtype=torch.int32, device="cuda")
    cu_k = torch.tensor(np.concatenate([[0], np.cumsum(hist + 1)]), dtype=torch.int32, device="cuda")
    d_tok = torch.from_numpy(tokens).cuda()
    d_out = torch.zeros(B, dtype=torch.int32, device="cuda")
    m.decode_step(batch=B, token_ids=d_tok, cu_q_len=cu_q, cu_k_len=cu_k, block_ptrs=ptrs, cu_block_nums=cu_blocks,
                  out_tokens=d_out, max_k_len=int(hist.max()) + 1)
    torch.cuda.synchronize()
    got_logits = from_dev_T(m.logits(B), BF16)
    got_tokens = d_out.cpu().numpy()
    ref_logits, ref_tokens = decode_step(w, tokens, cache, tables, hist, rope)
    tol = 3.2e-2
    err = np.abs(got_logits - ref_logits)
    assert err.max() <= tol, err.max()
    srt = np.sort(ref_logits, -1)
    margin = srt[:, -1] - srt[:, -2]
    for b in range(B):
        if margin[b] > 2 * tol:
            assert got_tokens[b] == ref_tokens[b]
        else:
            assert ref_logits[b, got_tokens[b]] >= srt[b, -1] - 2 * tol


def test_synthetic_fp8_model_runs_and_is_deterministic():
    """DecodeModel.load_synthetic for an fp8 spec (the bench path of config C5): finite logits, identical across runs."""
    import torch
    from lmdeploy_b200.turbomind import DecodeModel, KvCachePool, ModelSpec
    spec = ModelSpec("tiny-fp8", "llama", 1024, 2, 8, 2, 2048, 2048, 5e5, weight_format="fp8", dtype="bf16")
    outs = []
    for _ in range(2):
        m = DecodeModel(spec, quant_policy=8, max_batch=4)
        m.load_synthetic(0)
        pool = KvCachePool(8, spec.layers, 2, 128, 64, 8)
        tables = [pool.allocate(2) for _ in range(4)]
        ptrs = torch.tensor([pool.ptr(b) for t in tables for b in t], dtype=torch.int64, device="cuda")
        cu_blocks = torch.arange(5, dtype=torch.int32, device="cuda") * 2
        cu = torch.arange(5, dtype=torch.int32, device="cuda")
        tok = torch.tensor([1, 2, 3, 4], dtype=torch.int32, device="cuda")
        out = torch.zeros(4, dtype=torch.int32, device="cuda")
        m.decode_step(batch=4, token_ids=tok, cu_q_len=cu, cu_k_len=cu.clone(), block_ptrs=ptrs, cu_block_nums=cu_blocks,
                      out_tokens=out, max_k_len=1)
        torch.cuda.synchronize()
        lg = m.logits(4).float().cpu()
        assert torch.isfinite(lg).all() and lg.abs().max() > 0
        outs.append((lg, out.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


# ---------------------------------------------------------------------------------------------------------------
# Qwen2 family: q/k/v bias applied by the attention prologue inside the decode step (B200_T_QKV_BIAS)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["awq", "fp16"])
def test_decode_step_with_qkv_bias_matches_oracle(fmt):
    import torch
    from oracle import attention as OA
    from oracle.decode_step import TinyWeights, decode_step
    from harness import from_dev_T
    from test_decode_step_gpu import build_device_model
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=1024, vocab=1024, fmt=fmt, seed=13, qkv_bias=True)
    hist = np.array([0, 7, 64, 200], np.int32)
    B = len(hist)
    rng = np.random.default_rng(4)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)     # build_device_model's theta
    layout = OA.KvLayout(2, 64, 128, 8)
    blocks_per = [(int(h) + 1 + 63) // 64 for h in hist]
    cache = OA.PagedKvCache(sum(blocks_per) + 1, w.layers, layout, F16)
    tables, o = [], 0
    for n in blocks_per:
        tables.append(list(range(o, o + n)))
        o += n
    for l in range(w.layers):
        for b in range(B):
            if hist[b]:
                OA.process_kv(cache, tables[b], l, 0, OA.rn(rng.standard_normal((int(hist[b]), 2, 128)).astype(np.float32), F16),
                              OA.rn(rng.standard_normal((int(hist[b]), 2, 128)).astype(np.float32), F16), rope)
    tokens = rng.integers(0, w.vocab, B).astype(np.int32)
    m = build_device_model(w, 8, max_batch=8)
    pool = torch.from_numpy(cache.pool.copy()).cuda()
    ptrs = torch.tensor([pool.data_ptr() + int(x) * cache.block_bytes for t in tables for x in t], dtype=torch.int64, device="cuda")
    cu_blocks = torch.tensor(np.concatenate([[0], np.cumsum(blocks_per)]), dtype=torch.int32, device="cuda")
    cu_q = torch.arange(B + 1, dtype=torch.int32, device="cuda")
    cu_k = torch.tensor(np.concatenate([[0], np.cumsum(hist + 1)]), dtype=torch.int32, device="cuda")
    d_out = torch.zeros(B, dtype=torch.int32, device="cuda")
    m.decode_step(batch=B, token_ids=torch.from_numpy(tokens).cuda(), cu_q_len=cu_q, cu_k_len=cu_k, block_ptrs=ptrs,
                  cu_block_nums=cu_blocks, out_tokens=d_out, max_k_len=int(hist.max()) + 1)
    torch.cuda.synchronize()
    got = from_dev_T(m.logits(B), F16)
    ref, _ = decode_step(w, tokens, cache, tables, hist, rope)
    assert np.abs(got - ref).max() <= 4e-3
    assert (pool.cpu().numpy() == cache.pool).mean() > 0.999


@pytest.mark.parametrize("heads", [(14, 2), (6, 2), (5, 1)])
def test_decode_attention_odd_gqa_groups(heads):
    """GQA groups 7 (Qwen2-7B), 3 and 5: the decode kernel pads the group to the MMA's 8 columns; the verified suite
    covers groups 1, 4 and 8."""
    from harness import PagedCase
    hq, hkv = heads
    case = PagedCase(batch_hist=[0, 63, 64, 300, 1000], num_heads=hq, num_kv_heads=hkv, kv_bits=8, seed=hq, bias=True)
    out, pool = case.run_device()
    ref = case.run_oracle()
    assert (pool == case.cache.pool).all()
    assert np.abs(out - ref).max() <= 2e-3


# ---------------------------------------------------------------------------------------------------------------
# Qwen3 family: per-head q / k RMSNorm (b200_qk_rmsnorm, csrc/qk_norm.cu) and the decode step that uses it
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [F16, BF16])
def test_qk_rmsnorm_matches_oracle(lib, dtype):
    """The kernel reproduces the reference kernel's instruction order (FMA chain, butterfly, fma(sum, 1/dim, eps), rsqrt
    approximation); the numpy restatement has an exact rsqrt, so an output may sit one T ulp away -- rarely."""
    import ctypes as C
    import torch
    from lmdeploy_b200 import _lib as L
    from harness import to_dev_T, from_dev_T
    rng = np.random.default_rng(9)
    Hq, Hkv, T = 8, 2, 37
    width = (Hq + 2 * Hkv) * 128
    qkv = rn((rng.standard_normal((T, width)) * 2).astype(np.float32), dtype)
    qw = rn((1 + 0.2 * rng.standard_normal(128)).astype(np.float32), dtype)
    kw = rn((1 + 0.2 * rng.standard_normal(128)).astype(np.float32), dtype)
    d = to_dev_T(qkv, dtype)
    dq, dk = to_dev_T(qw, dtype), to_dev_T(kw, dtype)
    L.check(lib.b200_qk_rmsnorm(d.data_ptr(), width, dq.data_ptr(), dk.data_ptr(), T, Hq, Hkv, 128, 1e-6,
                                0 if dtype == F16 else 1, torch.cuda.current_stream().cuda_stream), "b200_qk_rmsnorm")
    torch.cuda.synchronize()
    got = from_dev_T(d, dtype)
    want = OL.qk_rmsnorm(qkv, qw, kw, Hq, Hkv, 128, 1e-6, dtype)
    assert np.array_equal(got[:, (Hq + Hkv) * 128:], qkv[:, (Hq + Hkv) * 128:])           # v untouched
    ulp = np.abs(want) * (2.0 ** -10 if dtype == F16 else 2.0 ** -7) + 1e-6
    assert (np.abs(got - want) <= ulp).all()
    assert (got != want).mean() < 0.02


@pytest.mark.parametrize("fmt", ["awq", "fp16"])
def test_decode_step_with_qk_norm_matches_oracle(fmt):
    import torch
    from oracle import attention as OA
    from oracle.decode_step import TinyWeights, decode_step
    from test_decode_step_gpu import build_device_model
    from harness import from_dev_T
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=1024, vocab=1024, fmt=fmt, seed=17, qk_norm=True)
    hist = np.array([0, 5, 64, 130], np.int32)
    B = len(hist)
    rng = np.random.default_rng(4)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    layout = OA.KvLayout(2, 64, 128, 8)
    blocks_per = [(int(h) + 1 + 63) // 64 for h in hist]
    nblk = sum(blocks_per) + 2
    cache = OA.PagedKvCache(nblk, w.layers, layout, F16)
    perm = rng.permutation(nblk)
    tables, o = [], 0
    for n in blocks_per:
        tables.append(perm[o:o + n].tolist())
        o += n
    for l in range(w.layers):
        for b in range(B):
            h = int(hist[b])
            if h:
                OA.process_kv(cache, tables[b], l, 0, rn(rng.standard_normal((h, 2, 128)).astype(np.float32), F16),
                              rn(rng.standard_normal((h, 2, 128)).astype(np.float32), F16), rope)
    tokens = rng.integers(0, w.vocab, B).astype(np.int32)
    m = build_device_model(w, 8, max_batch=8)
    pool = torch.from_numpy(cache.pool.copy()).cuda()
    ptrs = torch.tensor([pool.data_ptr() + int(b) * cache.block_bytes for t in tables for b in t], dtype=torch.int64,
                        device="cuda")
    cu_blocks = torch.tensor(np.concatenate([[0], np.cumsum(blocks_per)]), dtype=torch.int32, device="cuda")
    cu_q = torch.arange(B + 1, dtype=torch.int32, device="cuda")
    cu_k = torch.tensor(np.concatenate([[0], np.cumsum(hist + 1)]), dtype=torch.int32, device="cuda")
    d_out = torch.zeros(B, dtype=torch.int32, device="cuda")
    m.decode_step(batch=B, token_ids=torch.from_numpy(tokens).cuda(), cu_q_len=cu_q, cu_k_len=cu_k, block_ptrs=ptrs,
                  cu_block_nums=cu_blocks, out_tokens=d_out, max_k_len=int(hist.max()) + 1)
    torch.cuda.synchronize()
    got = from_dev_T(m.logits(B), F16)
    ref, ref_tokens = decode_step(w, tokens, cache, tables, hist, rope, eps=m.cfg.norm_eps)
    assert np.abs(got - ref).max() <= 4e-3
    srt = np.sort(ref, -1)
    for b in range(B):
        if srt[b, -1] - srt[b, -2] > 8e-3:
            assert d_out.cpu().numpy()[b] == ref_tokens[b]


@pytest.mark.parametrize("dtype", [F16, BF16])
def test_qk_rmsnorm_is_bit_identical_to_the_reference_kernel(lib, dtype):
    """Ours against the REFERENCE's invokeQkRMSNorm (oracle/_ref/libtm_ref_norm.so) on the same buffers.  The kernel was
    written to the reference's compiled instruction order, so the bar is equality, not a tolerance."""
    import torch
    from lmdeploy_b200 import _lib as L
    from oracle import ref_lib
    from harness import to_dev_T
    if not ref_lib.norm_available():
        pytest.skip("oracle/_ref/libtm_ref_norm.so not built")
    ref = ref_lib.load_norm()
    if not hasattr(ref, "ref_qk_rmsnorm"):
        pytest.skip("libtm_ref_norm.so predates ref_qk_rmsnorm")
    rng = np.random.default_rng(12)
    for (Hq, Hkv, T) in ((8, 2, 37), (32, 8, 64), (28, 4, 5)):
        width = (Hq + 2 * Hkv) * 128
        qkv = rn((rng.standard_normal((T, width)) * rng.uniform(0.01, 30)).astype(np.float32), dtype)
        qw = rn((1 + 0.2 * rng.standard_normal(128)).astype(np.float32), dtype)
        kw = rn((1 + 0.2 * rng.standard_normal(128)).astype(np.float32), dtype)
        a, b = to_dev_T(qkv, dtype), to_dev_T(qkv, dtype)
        dq, dk = to_dev_T(qw, dtype), to_dev_T(kw, dtype)
        st = torch.cuda.current_stream().cuda_stream
        dt = 0 if dtype == F16 else 1
        L.check(lib.b200_qk_rmsnorm(a.data_ptr(), width, dq.data_ptr(), dk.data_ptr(), T, Hq, Hkv, 128, 1e-6, dt, st), "ours")
        assert ref.ref_qk_rmsnorm(b.data_ptr(), width, dq.data_ptr(), dk.data_ptr(), T, Hq, Hkv, 128, 1e-6, dt, st) == 0
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), (Hq, Hkv, T)
