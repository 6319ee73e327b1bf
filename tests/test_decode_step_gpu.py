"""GPU parity of the whole decode step (b200_model_decode_step) against the oracle's decode_step on a tiny
llama-family model with real operator shapes (head_dim 128, u4 g128 weights, KV int8/int4/fp16).

Bars: logits |diff| <= 4e-3 (three chained fp16 GEMM layers; the north star's single-op bar is 1e-3) and
greedy token ids identical wherever the oracle's top-1/top-2 logit margin exceeds that tolerance
(exact fp16 ties are legitimately ambiguous)."""
import ctypes as C

import numpy as np
import pytest

from oracle import attention as OA
from oracle import linear as OL
from oracle.decode_step import TinyWeights, decode_step
from oracle.numerics import F16

from harness import to_dev_T, from_dev_T

pytestmark = pytest.mark.gpu


def build_device_model(w: TinyWeights, quant_policy, max_batch, tp_size=1, tp_rank=0, shard=None, rope_exact_math=True):
    import torch
    from lmdeploy_b200.turbomind import ModelSpec, DecodeModel
    T = w.dtype
    spec = ModelSpec("tiny", "llama", w.hidden, w.layers, w.heads, w.kv_heads, w.inter, w.vocab, 1e4,
                     weight_format=w.fmt, dtype="f16" if T == F16 else "bf16", qkv_bias="qkv_bias" in w.L[0],
                     qk_norm="q_norm" in w.L[0])
    m = DecodeModel(spec, quant_policy=quant_policy, max_batch=max_batch, tp_size=tp_size, tp_rank=tp_rank,
                    rope_exact_math=rope_exact_math)   # True: the numpy oracle cannot reproduce MUFU approximations
    sw = shard if shard is not None else w

    def lin(d):
        if w.fmt == "awq":
            return dict(weight=torch.from_numpy(OL.pack_u4_row(d["q"])).cuda(), scales=to_dev_T(d["scales"], T),
                        zeros=to_dev_T(d["zeros"], T))
        if w.fmt == "fp8":
            return dict(weight=torch.from_numpy(np.ascontiguousarray(d["w8"])).cuda(),
                        col_scales=torch.from_numpy(np.ascontiguousarray(d["scales"], np.float32)).cuda())
        return dict(weight=to_dev_T(np.ascontiguousarray(d["w"]), T))

    def provider(name):
        if name == "tok_embeddings":
            return to_dev_T(w.embed, T)
        if name == "norm":
            return to_dev_T(w.final_norm, T)
        if name == "output":
            return dict(weight=to_dev_T(np.ascontiguousarray(sw.output["w"]), T))
        parts = name.split(".")
        i, kind = int(parts[1]), parts[2]
        if kind == "attention_norm":
            return to_dev_T(w.L[i]["attn_norm"], T)
        if kind == "ffn_norm":
            return to_dev_T(w.L[i]["ffn_norm"], T)
        if kind == "qkv_bias":
            return to_dev_T(sw.L[i]["qkv_bias"], T)
        if kind in ("q_norm", "k_norm"):
            return to_dev_T(sw.L[i][kind], T)
        return lin(sw.L[i][kind])

    m.load_tm_weights(provider)
    return m


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
@pytest.mark.parametrize("fmt", ["awq", "fp16"])
def test_decode_step_matches_oracle(kv_bits, fmt):
    import torch
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=1024, vocab=1024, fmt=fmt, seed=kv_bits)
    hist = np.array([0, 5, 64, 130, 311], np.int32)
    B = len(hist)
    rng = np.random.default_rng(1)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    layout = OA.KvLayout(2, 64, 128, kv_bits)
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
                k = OA.rn(rng.standard_normal((h, 2, 128)).astype(np.float32), F16)
                v = OA.rn(rng.standard_normal((h, 2, 128)).astype(np.float32), F16)
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
    got_logits = from_dev_T(m.logits(B), F16)
    got_tokens = d_out.cpu().numpy()

    ref_logits, ref_tokens = decode_step(w, tokens, cache, tables, hist, rope)
    # KV pages: layer 0 is bit-exact; deeper layers see GEMM outputs that may differ from the oracle by one fp16
    # ulp (fp32 accumulation order), which moves a few quantised bytes / raw fp16 values by one step.
    same = (pool.cpu().numpy() == cache.pool).mean()
    assert same > (0.999 if kv_bits != 16 else 0.995), same
    err = np.abs(got_logits - ref_logits)
    assert err.max() <= 4e-3, err.max()
    srt = np.sort(ref_logits, -1)
    margin = srt[:, -1] - srt[:, -2]
    for b in range(B):
        if margin[b] > 8e-3:
            assert got_tokens[b] == ref_tokens[b]
        else:
            assert ref_logits[b, got_tokens[b]] >= srt[b, -1] - 8e-3


def test_multi_step_advance_and_determinism():
    """advance=1 feeds sampled tokens back on the device; two identical runs give identical token streams."""
    import torch
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=1024, vocab=1024, seed=3)
    B, steps = 4, 70       # crosses a 64-token block boundary
    streams = []
    for rep in range(2):
        m = build_device_model(w, 8, max_batch=B)
        from lmdeploy_b200.turbomind import KvCachePool
        pool = KvCachePool(B * 3, w.layers, 2, 128, 64, 8)
        tables = [pool.allocate(3) for _ in range(B)]
        ptrs = torch.tensor([pool.ptr(b) for t in tables for b in t], dtype=torch.int64, device="cuda")
        cu_blocks = torch.arange(B + 1, dtype=torch.int32, device="cuda") * 3
        cu_q = torch.arange(B + 1, dtype=torch.int32, device="cuda")
        cu_k = torch.arange(B + 1, dtype=torch.int32, device="cuda")           # k_len = 1 each
        tok = torch.tensor([1, 2, 3, 4], dtype=torch.int32, device="cuda")
        out = torch.zeros(B, dtype=torch.int32, device="cuda")
        hist = []
        for s in range(steps):
            m.decode_step(batch=B, token_ids=tok, cu_q_len=cu_q, cu_k_len=cu_k, block_ptrs=ptrs,
                          cu_block_nums=cu_blocks, out_tokens=out, max_k_len=s + 1, advance=True)
            hist.append(out.clone())
        torch.cuda.synchronize()
        assert cu_k.cpu().tolist() == [i * (steps + 1) for i in range(B + 1)]
        streams.append(torch.stack(hist).cpu().numpy())
    assert (streams[0] == streams[1]).all()
    assert streams[0].min() >= 0 and streams[0].max() < w.vocab
