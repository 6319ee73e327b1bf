"""Oracle self-consistency for the prefill restatement: attending a chunk in one call must equal feeding the same tokens
through the (golden-pinned) decode restatement one at a time -- outputs and the KV pages they leave behind, exactly."""
import numpy as np
import pytest

from oracle import attention as OA
from oracle.numerics import BF16, F16, rn


@pytest.mark.parametrize("kv_bits,dtype", [(8, F16), (4, F16), (16, F16), (8, BF16)])
def test_prefill_chunk_equals_token_by_token_decode(kv_bits, dtype):
    rng = np.random.default_rng(kv_bits)
    Hq, Hkv, D = 4, 2, 128
    lay = OA.KvLayout(Hkv, 64, D, kv_bits)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=D, base=1e4)
    q_lens, hists = [5, 3], [70, 0]
    tables = [[2, 0], [1]]
    caches = [OA.PagedKvCache(4, 1, lay, dtype) for _ in range(2)]
    for b, h in enumerate(hists):
        if h:
            k0 = rn(rng.standard_normal((h, Hkv, D)).astype(np.float32), dtype)
            v0 = rn(rng.standard_normal((h, Hkv, D)).astype(np.float32), dtype)
            for c in caches:
                OA.process_kv(c, tables[b], 0, 0, k0, v0, rope)
    qkv = rn(rng.standard_normal((sum(q_lens), (Hq + 2 * Hkv) * D)).astype(np.float32), dtype)
    bias = [rn((rng.standard_normal(n * D) * 0.1).astype(np.float32), dtype) for n in (Hq, Hkv, Hkv)]
    k_lens = [q + h for q, h in zip(q_lens, hists)]
    got = OA.prefill_attention(qkv, caches[0], tables, q_lens, k_lens, 0, Hq, Hkv, rope, q_bias=bias[0], k_bias=bias[1],
                               v_bias=bias[2])
    rows, t0 = [], 0
    for b, (ql, h) in enumerate(zip(q_lens, hists)):
        for i in range(ql):
            rows.append(OA.decode_attention(qkv[t0 + i:t0 + i + 1], caches[1], [tables[b]], np.array([h + i]), 0, Hq, Hkv,
                                            rope, q_bias=bias[0], k_bias=bias[1], v_bias=bias[2])[0])
        t0 += ql
    assert np.array_equal(got, np.stack(rows))
    assert np.array_equal(caches[0].pool, caches[1].pool)


def test_chunk_step_equals_token_by_token_decode_steps():
    """Model level: one mixed step (a 6-token prompt chunk + a decoding sequence) leaves the same pages and gives, row for
    row, the same logits as feeding those tokens through decode_step one at a time."""
    from oracle.decode_step import TinyWeights, chunk_step, decode_step
    w = TinyWeights(hidden=256, layers=2, heads=2, kv_heads=1, inter=256, vocab=256, fmt="awq", seed=4)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    lay = OA.KvLayout(1, 64, 128, 8)
    caches = [OA.PagedKvCache(4, w.layers, lay, F16) for _ in range(2)]
    rng = np.random.default_rng(0)
    # sequence 1 already has 70 cached tokens (fed one by one into both caches)
    hist_tokens = rng.integers(0, 256, 70)
    for c in caches:
        for i, t in enumerate(hist_tokens):
            decode_step(w, np.array([t]), c, [[1, 3]], np.array([i], np.int32), rope)
    prompt = rng.integers(0, 256, 6)
    nxt = np.array([17])
    tokens = np.concatenate([prompt, nxt])
    got, _ = chunk_step(w, tokens, caches[0], [[0], [1, 3]], [6, 1], [6, 71], rope)
    rows = []
    for i, t in enumerate(prompt):
        lg, _ = decode_step(w, np.array([t]), caches[1], [[0]], np.array([i], np.int32), rope)
        rows.append(lg[0])
    lg, _ = decode_step(w, nxt, caches[1], [[1, 3]], np.array([70], np.int32), rope)
    rows.append(lg[0])
    assert np.array_equal(got, np.stack(rows))
    assert np.array_equal(caches[0].pool, caches[1].pool)
