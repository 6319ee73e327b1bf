"""tcgen05 prefill attention + the mixed prefill/decode step: vs the oracle and vs the reference's own kernels
(ProcessKV + FlattenKV + dispatchAttention built in place, oracle/_ref/libtm_ref_prefill.so).  First device run: round 2."""
import numpy as np
import pytest

from oracle import linear as OL
from oracle.numerics import BF16, F16, rn

from test_linear_gpu import DevLinear, _assert_close

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------------
# prefill (context) attention: b200_prefill_attention vs oracle.attention.prefill_attention
# ---------------------------------------------------------------------------------------------------------------
def _prefill_case(lib, *, q_lens, hists, Hq, Hkv, kv_bits, dtype, seed=0, bias=False, rope=None, against_reference=False):
    import ctypes as C
    import torch
    from lmdeploy_b200 import _lib as L
    from oracle import attention as OA
    from harness import rand_T, to_dev_T, from_dev_T
    rng = np.random.default_rng(seed)
    D, block_len, layers, layer = 128, 64, 2, 1
    q_lens, hists = np.asarray(q_lens, np.int32), np.asarray(hists, np.int32)
    k_lens = q_lens + hists
    B = len(q_lens)
    rope = rope or OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=10000.0)
    layout = OA.KvLayout(head_num=Hkv, block_len=block_len, head_dim=D, q_bits=kv_bits)
    blocks_per = [(int(k) + block_len - 1) // block_len for k in k_lens]
    nblk = sum(blocks_per) + 2
    cache = OA.PagedKvCache(nblk, layers, layout, dtype)
    cache.pool[:] = rng.integers(0, 256, cache.pool.shape, dtype=np.uint8)      # pages start as garbage, not zeros
    perm = rng.permutation(nblk)
    tables, o = [], 0
    for n in blocks_per:
        tables.append(perm[o:o + n].tolist())
        o += n
    biases = None
    if bias:
        biases = (rand_T(rng, (Hq * D,), dtype, 0.1), rand_T(rng, (Hkv * D,), dtype, 0.1), rand_T(rng, (Hkv * D,), dtype, 0.1))
    for b in range(B):
        h = int(hists[b])
        if h:
            OA.process_kv(cache, tables[b], layer, 0, rand_T(rng, (h, Hkv, D), dtype), rand_T(rng, (h, Hkv, D), dtype), rope,
                          None if not bias else biases[1], None if not bias else biases[2])
    T = int(q_lens.sum())
    qkv = rand_T(rng, (T, (Hq + 2 * Hkv) * D), dtype)

    # ---- device ----
    d_pool = torch.from_numpy(cache.pool.copy()).cuda()
    ptrs = [d_pool.data_ptr() + int(x) * cache.block_bytes for t in tables for x in t]
    d_ptrs = torch.tensor(ptrs, dtype=torch.int64, device="cuda")
    d_cu_blocks = torch.tensor(np.concatenate([[0], np.cumsum(blocks_per)]), dtype=torch.int32, device="cuda")
    d_cu_q = torch.tensor(np.concatenate([[0], np.cumsum(q_lens)]), dtype=torch.int32, device="cuda")
    d_cu_k = torch.tensor(np.concatenate([[0], np.cumsum(k_lens)]), dtype=torch.int32, device="cuda")
    d_qkv = to_dev_T(qkv, dtype)
    d_out = torch.zeros(T, Hq * D, dtype=d_qkv.dtype, device="cuda")
    p = L.AttentionParams()
    base = d_qkv.data_ptr()
    p.out, p.q, p.k, p.v = d_out.data_ptr(), base, base + Hq * D * 2, base + (Hq + Hkv) * D * 2
    p.stride = (Hq + 2 * Hkv) * D
    if bias:
        d_bias = [to_dev_T(x, dtype) for x in biases]
        p.q_bias, p.k_bias, p.v_bias = [t.data_ptr() for t in d_bias]
    p.cu_q_len, p.cu_k_len = d_cu_q.data_ptr(), d_cu_k.data_ptr()
    p.block_ptrs, p.cu_block_nums = d_ptrs.data_ptr(), d_cu_blocks.data_ptr()
    p.layer_offset, p.block_len = cache.layer_offset(layer), block_len
    p.token_num, p.batch_size, p.max_k_len = T, B, int(k_lens.max())
    p.num_heads, p.num_kv_heads, p.size_per_head = Hq, Hkv, D
    p.inv_sqrt_dh = float(np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D)))))
    p.rope = L.RopeConfig(rope.type, rope.dim, rope.base, rope.factor, rope.max_position_embeddings,
                          rope.llama3_low_freq_factor, rope.llama3_high_freq_factor,
                          rope.llama3_original_max_position_embeddings, rope.yarn_attention_factor, rope.yarn_beta_fast,
                          rope.yarn_beta_slow, 1)
    p.rope.exact_math = 0 if against_reference else 1
    p.quant_policy, p.dtype = (0 if kv_bits == 16 else kv_bits), (0 if dtype == F16 else 1)
    ws = torch.zeros(lib.b200_prefill_attention_workspace_bytes(T, Hq), dtype=torch.uint8, device="cuda")
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    p.stream = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200_prefill_attention(C.byref(p)), "b200_prefill_attention")
    torch.cuda.synchronize()
    got, got_pool = from_dev_T(d_out, dtype), d_pool.cpu().numpy()

    if against_reference:
        # the reference's ProcessKV + FlattenKV + dispatchAttention on a fresh copy of the same pages
        from oracle import ref_lib
        ref = ref_lib.load_prefill()
        d_pool2 = torch.from_numpy(cache.pool.copy()).cuda()
        d_ptrs2 = torch.tensor([d_pool2.data_ptr() + int(x) * cache.block_bytes for t in tables for x in t],
                               dtype=torch.int64, device="cuda")
        d_out2 = torch.zeros_like(d_out)
        p.block_ptrs, p.out = d_ptrs2.data_ptr(), d_out2.data_ptr()
        a, keep = ref_lib.args_from_b200(p, torch, max_q_len=int(q_lens.max()))
        k_sum = int(k_lens.sum())
        tmp_kv = torch.zeros(Hkv * 2 * (k_sum + 64) * D, dtype=d_qkv.dtype, device="cuda")
        rc = ref.ref_prefill_attention(C.byref(a), tmp_kv.data_ptr(), k_sum)
        torch.cuda.synchronize()
        assert rc == 0
        return got, from_dev_T(d_out2, dtype), got_pool, d_pool2.cpu().numpy()

    # ---- oracle (mutates its cache the same way) ----
    ref = OA.prefill_attention(qkv, cache, tables, q_lens, k_lens, layer, Hq, Hkv, rope,
                               q_bias=None if not bias else biases[0], k_bias=None if not bias else biases[1],
                               v_bias=None if not bias else biases[2])
    return got, ref, got_pool, cache.pool


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
@pytest.mark.parametrize("heads", [(4, 1), (8, 2), (2, 2), (6, 1)])          # GQA groups 4, 4, 1 (MHA), 6 (126 rows used)
def test_prefill_attention_matches_oracle(lib, kv_bits, heads):
    """Ragged chunks: fresh prompts, chunked continuation over history, single-token chunk, chunk ending mid-tile and
    mid-block, more query tokens than one 128-row tile holds.  Bars as for decode: KV pages bit-identical (the prologue
    kernel is shared), outputs |diff| <= 2e-3 (fp16; tile order differs from the oracle's single pass)."""
    Hq, Hkv = heads
    got, ref, pool, ref_pool = _prefill_case(lib, q_lens=[70, 1, 33, 200, 5], hists=[0, 130, 64, 0, 63], Hq=Hq, Hkv=Hkv,
                                             kv_bits=kv_bits, dtype=F16, seed=kv_bits + Hq)
    assert np.array_equal(pool, ref_pool)
    assert np.isfinite(got).all()
    err = np.abs(got - ref)
    assert err.max() <= 2e-3, (err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("kv_bits", [8, 16])
def test_prefill_attention_bf16_bias_llama3_rope(lib, kv_bits):
    from oracle import attention as OA
    rope = OA.RopeParam(type=OA.ROPE_LLAMA3, dim=128, base=5e5, factor=8.0, llama3_low_freq_factor=1.0,
                        llama3_high_freq_factor=4.0, llama3_original_max_position_embeddings=8192)
    got, ref, pool, ref_pool = _prefill_case(lib, q_lens=[150, 17], hists=[10, 300], Hq=8, Hkv=2, kv_bits=kv_bits,
                                             dtype=BF16, seed=3, bias=True, rope=rope)
    assert np.array_equal(pool, ref_pool)
    err = np.abs(got - ref)
    assert np.isfinite(got).all() and err.max() <= 1.6e-2, err.max()


def test_prefill_single_prompt_crossing_one_block(lib):
    """One 96-token prompt (two 64-token blocks, the second half full), no history."""
    got, ref, pool, ref_pool = _prefill_case(lib, q_lens=[96], hists=[0], Hq=8, Hkv=2, kv_bits=8, dtype=F16, seed=11)
    assert np.array_equal(pool, ref_pool) and np.abs(got - ref).max() <= 2e-3


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
def test_prefill_attention_matches_reference_kernels(lib, kv_bits):
    """Ours against the REFERENCE ITSELF (oracle/_ref/libtm_ref_prefill.so: invokeProcessKV_v2 + invokeFlattenKV_v2 +
    dispatchAttention compiled in place) on the same inputs, rope.exact_math = 0: pages bit-identical, outputs <= 2e-3."""
    from oracle import ref_lib
    if not ref_lib.prefill_available():
        pytest.skip("oracle/_ref/libtm_ref_prefill.so not built")
    got, ref, pool, ref_pool = _prefill_case(lib, q_lens=[70, 1, 33, 200], hists=[0, 130, 64, 0], Hq=8, Hkv=2, kv_bits=kv_bits,
                                             dtype=F16, seed=kv_bits, against_reference=True)
    assert np.array_equal(pool, ref_pool)
    assert np.isfinite(got).all() and np.isfinite(ref).all()
    assert np.abs(got - ref).max() <= 2e-3


# ---------------------------------------------------------------------------------------------------------------
# model level: b200_model_chunk_step (mixed prefill/decode rows) vs oracle.decode_step.chunk_step
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kv_bits", [8, 16])
def test_chunk_step_matches_oracle(kv_bits):
    """Three sequences in one step: a fresh 40-token prompt chunk, a decoding sequence with 130 cached tokens, a 9-token
    continuation chunk over 64 cached tokens.  Bars as for the decode step: logits of EVERY row |diff| <= 4e-3, greedy ids
    equal outside top-1/top-2 ties; pages >= 99.9 % identical bytes (deeper layers see GEMM outputs one ulp apart)."""
    import torch
    from oracle import attention as OA
    from oracle.decode_step import TinyWeights, chunk_step
    from harness import from_dev_T
    from test_decode_step_gpu import build_device_model
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=1024, vocab=1024, fmt="awq", seed=kv_bits)
    q_lens, hists = np.array([40, 1, 9], np.int32), np.array([0, 130, 64], np.int32)
    k_lens = q_lens + hists
    B, T = len(q_lens), int(q_lens.sum())
    rng = np.random.default_rng(2)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    layout = OA.KvLayout(2, 64, 128, kv_bits)
    blocks_per = [(int(k) + 63) // 64 for k in k_lens]
    nblk = sum(blocks_per) + 2
    cache = OA.PagedKvCache(nblk, w.layers, layout, F16)
    perm = rng.permutation(nblk)
    tables, o = [], 0
    for n in blocks_per:
        tables.append(perm[o:o + n].tolist())
        o += n
    for l in range(w.layers):
        for b in range(B):
            h = int(hists[b])
            if h:
                OA.process_kv(cache, tables[b], l, 0, OA.rn(rng.standard_normal((h, 2, 128)).astype(np.float32), F16),
                              OA.rn(rng.standard_normal((h, 2, 128)).astype(np.float32), F16), rope)
    tokens = rng.integers(0, w.vocab, T).astype(np.int32)

    m = build_device_model(w, 0 if kv_bits == 16 else kv_bits, max_batch=64)
    pool = torch.from_numpy(cache.pool.copy()).cuda()
    ptrs = torch.tensor([pool.data_ptr() + int(x) * cache.block_bytes for t in tables for x in t], dtype=torch.int64,
                        device="cuda")
    cu_blocks = torch.tensor(np.concatenate([[0], np.cumsum(blocks_per)]), dtype=torch.int32, device="cuda")
    cu_q = torch.tensor(np.concatenate([[0], np.cumsum(q_lens)]), dtype=torch.int32, device="cuda")
    cu_k = torch.tensor(np.concatenate([[0], np.cumsum(k_lens)]), dtype=torch.int32, device="cuda")
    d_tok = torch.from_numpy(tokens).cuda()
    d_out = torch.zeros(T, dtype=torch.int32, device="cuda")
    m.decode_step(batch=B, token_ids=d_tok, cu_q_len=cu_q, cu_k_len=cu_k, block_ptrs=ptrs, cu_block_nums=cu_blocks,
                  out_tokens=d_out, max_k_len=int(k_lens.max()), token_num=T)
    torch.cuda.synchronize()
    got_logits = from_dev_T(m.logits(T), F16)
    got_tokens = d_out.cpu().numpy()
    ref_logits, ref_tokens = chunk_step(w, tokens, cache, tables, q_lens, k_lens, rope)
    same = (pool.cpu().numpy() == cache.pool).mean()
    assert same > (0.999 if kv_bits != 16 else 0.995), same
    err = np.abs(got_logits - ref_logits)
    assert err.max() <= 4e-3, (err.max(), np.unravel_index(err.argmax(), err.shape))
    srt = np.sort(ref_logits, -1)
    for t in range(T):
        if srt[t, -1] - srt[t, -2] > 8e-3:
            assert got_tokens[t] == ref_tokens[t]
