"""CPU oracle: one full decode step of a TurboMind llama-family model (test infrastructure only).

Layer order follows UnifiedDecoder::Forward (src/turbomind/models/llama/unified_decoder.cc:226-346):
  x = RMSNorm(residual; attention_norm[0])
  per layer:  qkv = x.Wqkv ; attn = decode_attention(qkv) ; o = attn.Wo
              residual += o ; x = RMSNorm(residual; ffn_norm)            (AllreduceResidualRMSnorm #1)
              a = gated_silu(x.W13) ; f = a.W2
              residual += f ; x = RMSNorm(residual; next attention_norm | final norm)   (#2)
  logits = x.Wout ; token = argmax(float(logits))   (language_model.cc:285-335, generation.cc:270-276)
TP (tensor parallel) variant: heads / inter / vocab sharded, partial o / f summed over ranks in rank order in T
(comm/cuda_ipc/fused_allreduce.cu:25-167 pull variant), logits concatenated.
"""
import numpy as np

from . import attention as OA
from . import linear as OL
from .numerics import F16, rn, f32


class TinyWeights:
    """Random weights of a small llama-family model in TurboMind's loader-visible layout."""

    def __init__(self, *, hidden, layers, heads, kv_heads, inter, vocab, fmt="awq", dtype=F16, seed=0, head_dim=128,
                 qkv_bias=False, qk_norm=False):
        rng = np.random.default_rng(seed)
        self.hidden, self.layers, self.heads, self.kv_heads = hidden, layers, heads, kv_heads
        self.inter, self.vocab, self.fmt, self.dtype, self.D = inter, vocab, fmt, dtype, head_dim
        D = head_dim

        def lin(K, N):
            w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
            if fmt == "awq":
                q, s, z = OL.quantize_groupwise_u4(w, 128)
                return dict(q=q, scales=s, zeros=z, deq=OL.dequant_u4(q, s, z, dtype))
            if fmt == "fp8":
                # e4m3 with one f32 scale per (128-k group, column): amax / 448, the usual fp8 weight quantiser
                sc = (np.abs(w.reshape(K // 128, 128, N)).max(axis=1) / 448.0).astype(np.float32)
                w8 = OL.f32_to_e4m3(w / np.repeat(sc, 128, 0))
                return dict(w8=w8, scales=sc, deq=OL.dequant_fp8_cols(w8, sc, 128, dtype))
            w = rn(w, dtype)
            return dict(w=w, deq=w)

        self.embed = rn((rng.standard_normal((vocab, hidden)) * 0.5).astype(np.float32), dtype)
        self.final_norm = rn((1 + 0.02 * rng.standard_normal(hidden)).astype(np.float32), dtype)
        self.L = []
        for _ in range(layers):
            self.L.append(dict(
                attn_norm=rn((1 + 0.02 * rng.standard_normal(hidden)).astype(np.float32), dtype),
                ffn_norm=rn((1 + 0.02 * rng.standard_normal(hidden)).astype(np.float32), dtype),
                w_qkv=lin(hidden, (heads + 2 * kv_heads) * D), wo=lin(heads * D, hidden),
                w1w3=lin(hidden, 2 * inter), w2=lin(inter, hidden)))
            if qkv_bias:   # Qwen2 family: added to q/k/v by the attention prologue (unified_attention_layer.cc:495-497)
                self.L[-1]["qkv_bias"] = rn((rng.standard_normal((heads + 2 * kv_heads) * D) * 0.1).astype(np.float32), dtype)
        if qk_norm:        # Qwen3 family: per-head RMSNorm of q and k before RoPE (unified_attention_layer.cc:739)
            for lw in self.L:
                lw["q_norm"] = rn((1 + 0.1 * rng.standard_normal(D)).astype(np.float32), dtype)
                lw["k_norm"] = rn((1 + 0.1 * rng.standard_normal(D)).astype(np.float32), dtype)
        out = rn((rng.standard_normal((hidden, vocab)) * 0.1 / np.sqrt(hidden)).astype(np.float32), dtype)
        self.output = dict(w=out, deq=out)

    # ---- TP shard of the loader-visible tensors (builders/_base.py:102-113, attention.py:65-76, ffn.py:238-243)
    def shard(self, rank, tp):
        D, Hq, Hkv = self.D, self.heads // tp, self.kv_heads // tp
        I = self.inter // tp

        def cols(x, idx):
            return {k: (v[:, idx] if k in ("q", "deq", "w", "w8") else v[:, idx] if v.ndim == 2 else v) for k, v in x.items()}

        def rows(x, sl):
            out = {}
            for k, v in x.items():
                if k in ("q", "deq", "w", "w8"):
                    out[k] = v[sl]
                else:   # scales / zeros [K/128, N]
                    out[k] = v[sl.start // 128: sl.stop // 128]
            return out

        sh = TinyWeights.__new__(TinyWeights)
        sh.__dict__.update({k: v for k, v in self.__dict__.items() if k != "L"})
        sh.heads, sh.kv_heads, sh.inter = Hq, Hkv, I
        sh.vocab_local = self.vocab // tp
        sh.L = []
        for lw in self.L:
            qi = np.arange(rank * Hq * D, (rank + 1) * Hq * D)
            ki = self.heads * D + np.arange(rank * Hkv * D, (rank + 1) * Hkv * D)
            vi = (self.heads + self.kv_heads) * D + np.arange(rank * Hkv * D, (rank + 1) * Hkv * D)
            gu = np.arange(rank * 2 * I, (rank + 1) * 2 * I)      # interleaved (gate, up) pairs stay together
            extra = {"qkv_bias": lw["qkv_bias"][np.concatenate([qi, ki, vi])]} if "qkv_bias" in lw else {}
            extra.update({k: lw[k] for k in ("q_norm", "k_norm") if k in lw})      # per-head weights: replicated
            sh.L.append(dict(attn_norm=lw["attn_norm"], ffn_norm=lw["ffn_norm"], **extra,
                             w_qkv=cols(lw["w_qkv"], np.concatenate([qi, ki, vi])),
                             wo=rows(lw["wo"], slice(rank * Hq * D, (rank + 1) * Hq * D)),
                             w1w3=cols(lw["w1w3"], gu),
                             w2=rows(lw["w2"], slice(rank * I, (rank + 1) * I))))
        vs = slice(rank * sh.vocab_local, (rank + 1) * sh.vocab_local)
        sh.output = {k: v[:, vs] for k, v in self.output.items()}
        return sh


def _qkv(x, w, lw, eps, dtype):
    """QKV projection (+ the per-head q/k RMSNorm of the Qwen3 family)."""
    qkv = OL.linear(x, lw["w_qkv"]["deq"], dtype)
    if "q_norm" in lw:
        qkv = OL.qk_rmsnorm(qkv, lw["q_norm"], lw["k_norm"], w.heads, w.kv_heads, w.D, eps, dtype)
    return qkv


def _bias_args(w, lw):
    b = lw.get("qkv_bias")
    if b is None:
        return {}
    nq, nk = w.heads * w.D, w.kv_heads * w.D
    return dict(q_bias=b[:nq], k_bias=b[nq:nq + nk], v_bias=b[nq + nk:])


def decode_step(w: TinyWeights, token_ids, cache: OA.PagedKvCache, block_tables, hist, rope, eps=1e-5, tp_shards=None,
                caches=None):
    """Returns (logits [B, vocab] float32 exact in T, tokens [B]).  Mutates cache(s).
    tp_shards: list of sharded TinyWeights (one per rank) with per-rank `caches` -> TP semantics."""
    dtype = w.dtype
    residual = w.embed[np.asarray(token_ids)]
    shards = tp_shards if tp_shards else [w]
    kvc = caches if caches else [cache]
    x = OL.rmsnorm(residual, w.L[0]["attn_norm"], eps, dtype)
    for l in range(w.layers):
        parts = []
        for r, sw in enumerate(shards):
            lw = sw.L[l]
            qkv = _qkv(x, sw, lw, eps, dtype)
            attn = OA.decode_attention(qkv, kvc[r], block_tables, hist, l, sw.heads, sw.kv_heads, rope, **_bias_args(sw, lw))
            parts.append(OL.linear(attn, lw["wo"]["deq"], dtype))
        x, residual = OL.allreduce_residual_rmsnorm(parts, residual, w.L[l]["ffn_norm"], eps, dtype)
        parts = []
        for r, sw in enumerate(shards):
            lw = sw.L[l]
            a = OL.linear_gated_silu(x, lw["w1w3"]["deq"], dtype)
            parts.append(OL.linear(a, lw["w2"]["deq"], dtype))
        nxt = w.L[l + 1]["attn_norm"] if l + 1 < w.layers else w.final_norm
        x, residual = OL.allreduce_residual_rmsnorm(parts, residual, nxt, eps, dtype)
    logits = np.concatenate([OL.linear(x, sw.output["deq"], dtype) for sw in shards], axis=1)
    return logits, logits.argmax(-1)


def chunk_step(w: TinyWeights, token_ids, cache: OA.PagedKvCache, block_tables, q_lens, k_lens, rope, eps=1e-5):
    """Mixed prefill/decode step (b200_model_chunk_step; the reference's chunked prefill, unified_attention_layer.cc
    :613-651): sequence b contributes q_lens[b] consecutive rows, k_lens[b] counts them.  Every operator except attention
    is row-wise and identical to decode_step; attention is oracle.attention.prefill_attention.  Returns
    (logits [sum(q_lens), vocab], argmax per row); the caller samples from the last row of each sequence."""
    dtype = w.dtype
    residual = w.embed[np.asarray(token_ids)]
    x = OL.rmsnorm(residual, w.L[0]["attn_norm"], eps, dtype)
    for l in range(w.layers):
        lw = w.L[l]
        qkv = _qkv(x, w, lw, eps, dtype)
        attn = OA.prefill_attention(qkv, cache, block_tables, q_lens, k_lens, l, w.heads, w.kv_heads, rope,
                                    **_bias_args(w, lw))
        x, residual = OL.allreduce_residual_rmsnorm([OL.linear(attn, lw["wo"]["deq"], dtype)], residual,
                                                    lw["ffn_norm"], eps, dtype)
        a = OL.linear_gated_silu(x, lw["w1w3"]["deq"], dtype)
        nxt = w.L[l + 1]["attn_norm"] if l + 1 < w.layers else w.final_norm
        x, residual = OL.allreduce_residual_rmsnorm([OL.linear(a, lw["w2"]["deq"], dtype)], residual, nxt, eps, dtype)
    logits = OL.linear(x, w.output["deq"], dtype)
    return logits, logits.argmax(-1)
