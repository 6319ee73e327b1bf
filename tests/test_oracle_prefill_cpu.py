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
