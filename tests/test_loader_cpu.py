"""Loader transforms (HF checkpoint -> TurboMind pre-prepare layout) checked end to end on CPU:
an HF-semantics reference forward (separate q/k/v, half-split RoPE, gate/up, plain attention) on the ORIGINAL
checkpoint tensors must agree with the oracle's TurboMind-semantics decode step on the CONVERTED tensors."""
import os
import numpy as np
import pytest
import torch

from oracle import attention as OA
from oracle import linear as OL
from oracle.decode_step import TinyWeights, decode_step
from oracle.numerics import BF16, F16, rn

from lmdeploy_b200.turbomind import loader as LD

H, L, HQ, HKV, I, V, D = 512, 2, 4, 2, 512, 256, 128


def pack_plain_last(q):
    """uint8 [R, C] -> int32 [R, C/8], nibble i = element 8w+i of the last axis."""
    return OL.pack_u4_row(q)


def quantize_symmetric_u4(w, group=128):
    """GPTQ-sym / compressed-tensors style: zero point 8, q = clamp(rne(w/s) + 8, 0, 15)."""
    K, N = w.shape
    wg = w.reshape(K // group, group, N)
    s = np.maximum(np.abs(wg).max(1), 1e-5) / np.float32(7)
    s = rn(s.astype(np.float32), F16)
    q = np.clip(np.rint(wg / s[:, None]) + 8, 0, 15).astype(np.uint8).reshape(K, N)
    return q, s, np.full_like(s, 8)


def quantize_fp8_blocks(w):
    """128x128 block-scaled e4m3 (scale = amax / 448): -> (bytes [K, N], weight_scale_inv [K/128, N/128])."""
    K, N = w.shape
    blk = (np.abs(w.reshape(K // 128, 128, N // 128, 128)).max(axis=(1, 3)) / 448.0).astype(np.float32)
    return OL.f32_to_e4m3(w / np.repeat(np.repeat(blk, 128, 0), 128, 1)), blk


FAMILIES = (None, "awq", "gptq", "compressed-tensors", "fp8")


def make_checkpoint(arch, family, seed=0):
    """family: None (fp16) | 'awq' | 'gptq' | 'compressed-tensors'; True/False are accepted for 'awq'/None."""
    family = {True: "awq", False: None}.get(family, family)
    rng = np.random.default_rng(seed)
    t, deq = {}, {}

    def add(name, K, N):
        w = (rng.standard_normal((K, N)) * 0.3 / np.sqrt(K)).astype(np.float32)
        if family == "awq":
            q, s, z = OL.quantize_groupwise_u4(w, 128)
            t[name + ".qweight"] = torch.from_numpy(OL.awq_pack(q))
            t[name + ".qzeros"] = torch.from_numpy(OL.awq_pack(z.astype(np.uint8)))
            t[name + ".scales"] = torch.from_numpy(s.astype(np.float16))
            deq[name] = OL.dequant_u4(q, s, z, F16)
        elif family == "gptq":
            q, s, z = quantize_symmetric_u4(w)
            # qweight packs 8 consecutive K rows per word; qzeros packs (zero - 1) along N
            t[name + ".qweight"] = torch.from_numpy(pack_plain_last(np.ascontiguousarray(q.T)).T.copy())
            t[name + ".qzeros"] = torch.from_numpy(pack_plain_last((z - 1).astype(np.uint8)))
            t[name + ".scales"] = torch.from_numpy(s.astype(np.float16))
            t[name + ".g_idx"] = torch.arange(K, dtype=torch.int32) // 128
            deq[name] = OL.dequant_u4(q, s, z, F16)
        elif family == "compressed-tensors":
            q, s, z = quantize_symmetric_u4(w)
            t[name + ".weight_packed"] = torch.from_numpy(pack_plain_last(np.ascontiguousarray(q.T)))   # [N, K/8]
            t[name + ".weight_scale"] = torch.from_numpy(np.ascontiguousarray(s.T).astype(np.float16))  # [N, K/g]
            t[name + ".weight_shape"] = torch.tensor([N, K])
            deq[name] = OL.dequant_u4(q, s, z, F16)
        elif family == "fp8":
            w8, blk = quantize_fp8_blocks(w)
            t[name + ".weight"] = torch.from_numpy(np.ascontiguousarray(w8.T)).view(torch.float8_e4m3fn)   # [N, K]
            t[name + ".weight_scale_inv"] = torch.from_numpy(np.ascontiguousarray(blk.T))                  # [N/128, K/128]
            deq[name] = OL.dequant_fp8(w8, blk, 128, BF16)
        else:
            w = rn(w, F16)
            t[name + ".weight"] = torch.from_numpy(w.T.astype(np.float16).copy())      # HF: [out, in]
            deq[name] = w

    for i in range(L):
        if arch == "internlm2":
            p = f"model.layers.{i}."
            add(p + "attention.wqkv", H, (HQ + 2 * HKV) * D)
            add(p + "attention.wo", HQ * D, H)
            add(p + "feed_forward.w1", H, I), add(p + "feed_forward.w3", H, I), add(p + "feed_forward.w2", I, H)
            norms = (p + "attention_norm.weight", p + "ffn_norm.weight")
        else:
            p = f"model.layers.{i}."
            if arch == "qwen2":
                for nm, n in (("q_proj", HQ * D), ("k_proj", HKV * D), ("v_proj", HKV * D)):
                    t[p + f"self_attn.{nm}.bias"] = torch.from_numpy((rng.standard_normal(n) * 0.2).astype(np.float16))
            if arch == "qwen3":      # per-head RMSNorm weights of q and k, shared by every head
                for nm in ("q_norm", "k_norm"):
                    t[p + f"self_attn.{nm}.weight"] = torch.from_numpy((1 + 0.2 * rng.standard_normal(D)).astype(np.float16))
            add(p + "self_attn.q_proj", H, HQ * D), add(p + "self_attn.k_proj", H, HKV * D)
            add(p + "self_attn.v_proj", H, HKV * D), add(p + "self_attn.o_proj", HQ * D, H)
            add(p + "mlp.gate_proj", H, I), add(p + "mlp.up_proj", H, I), add(p + "mlp.down_proj", I, H)
            norms = (p + "input_layernorm.weight", p + "post_attention_layernorm.weight")
        for n in norms:
            t[n] = torch.from_numpy((1 + 0.05 * rng.standard_normal(H)).astype(np.float16))
    emb = "model.tok_embeddings.weight" if arch == "internlm2" else "model.embed_tokens.weight"
    head = "output.weight" if arch == "internlm2" else "lm_head.weight"
    t[emb] = torch.from_numpy((rng.standard_normal((V, H)) * (0.5 if arch != "qwen2" else 0.05)).astype(np.float16))
    t["model.norm.weight"] = torch.from_numpy((1 + 0.05 * rng.standard_normal(H)).astype(np.float16))
    if arch != "qwen2":          # the Qwen2 test checkpoint ties lm_head to the embeddings (tie_word_embeddings)
        t[head] = torch.from_numpy((rng.standard_normal((V, H)) * 0.3 / np.sqrt(H)).astype(np.float16))
    cfg = {"architectures": [{"internlm2": "InternLM2ForCausalLM", "llama": "LlamaForCausalLM",
                              "qwen2": "Qwen2ForCausalLM", "qwen3": "Qwen3ForCausalLM"}[arch]],
           "tie_word_embeddings": arch == "qwen2", "head_dim": D,
           "hidden_size": H, "num_hidden_layers": L, "num_attention_heads": HQ, "num_key_value_heads": HKV,
           "intermediate_size": I, "vocab_size": V, "rms_norm_eps": 1e-5, "rope_theta": 10000.0}
    if family == "awq":
        cfg["quantization_config"] = {"quant_method": "awq", "bits": 4, "group_size": 128, "zero_point": True,
                                      "version": "gemm"}
    elif family == "gptq":
        cfg["quantization_config"] = {"quant_method": "gptq", "bits": 4, "group_size": 128, "sym": True,
                                      "desc_act": False}
    elif family == "fp8":
        cfg["quantization_config"] = {"quant_method": "fp8", "fmt": "e4m3", "weight_block_size": [128, 128]}
        cfg["torch_dtype"] = "bfloat16"
    elif family == "compressed-tensors":
        cfg["quantization_config"] = {"quant_method": "compressed-tensors", "format": "pack-quantized",
                                      "config_groups": {"group_0": {"format": "pack-quantized", "targets": ["Linear"],
                                                                    "weights": {"num_bits": 4, "type": "int",
                                                                                "group_size": 128, "symmetric": True,
                                                                                "strategy": "group"}}}}
    return cfg, t, deq


def hf_forward(arch, t, deq, tokens):
    """float64 HF-semantics forward over the whole token list; returns logits of every position."""
    f = lambda n: t[n].to(torch.float64).numpy()
    emb = f("model.tok_embeddings.weight" if arch == "internlm2" else "model.embed_tokens.weight")
    x = emb[np.asarray(tokens)]
    S = len(tokens)

    def norm(x, w):
        return x / np.sqrt((x ** 2).mean(-1, keepdims=True) + 1e-5) * w

    inv = 1.0 / (10000.0 ** (np.arange(0, D, 2) / D))
    ang = np.arange(S)[:, None] * inv[None]
    cos, sin = np.cos(ang), np.sin(ang)

    def rope(v):           # v [S, heads, D], HF half-split rotation
        a, b = v[..., : D // 2], v[..., D // 2:]
        return np.concatenate([a * cos[:, None] - b * sin[:, None], b * cos[:, None] + a * sin[:, None]], -1)

    for i in range(L):
        p = f"model.layers.{i}."
        if arch == "internlm2":
            h = norm(x, f(p + "attention_norm.weight"))
            qkv = (h @ deq[p + "attention.wqkv"]).reshape(S, HKV, HQ // HKV + 2, D)
            q = qkv[:, :, : HQ // HKV].reshape(S, HQ, D)
            k, v = qkv[:, :, HQ // HKV], qkv[:, :, HQ // HKV + 1]
            wo, w1, w3, w2 = (deq[p + "attention.wo"], deq[p + "feed_forward.w1"], deq[p + "feed_forward.w3"],
                              deq[p + "feed_forward.w2"])
            n2 = f(p + "ffn_norm.weight")
        else:
            h = norm(x, f(p + "input_layernorm.weight"))
            bq, bk, bv = [f(p + f"self_attn.{nm}.bias") if p + f"self_attn.{nm}.bias" in t else 0.0
                          for nm in ("q_proj", "k_proj", "v_proj")]
            q = (h @ deq[p + "self_attn.q_proj"] + bq).reshape(S, HQ, D)
            k = (h @ deq[p + "self_attn.k_proj"] + bk).reshape(S, HKV, D)
            v = (h @ deq[p + "self_attn.v_proj"] + bv).reshape(S, HKV, D)
            wo, w1, w3, w2 = (deq[p + "self_attn.o_proj"], deq[p + "mlp.gate_proj"], deq[p + "mlp.up_proj"],
                              deq[p + "mlp.down_proj"])
            n2 = f(p + "post_attention_layernorm.weight")
        if arch == "qwen3":      # HF Qwen3Attention: q_norm / k_norm over head_dim, then RoPE
            q = norm(q, f(p + "self_attn.q_norm.weight"))
            k = norm(k, f(p + "self_attn.k_norm.weight"))
        q, k = rope(q), rope(k)
        g = HQ // HKV
        out = np.zeros((S, HQ, D))
        for hq in range(HQ):
            s = q[:, hq] @ k[:, hq // g].T / np.sqrt(D)
            s = np.where(np.tril(np.ones((S, S))) > 0, s, -np.inf)
            pr = np.exp(s - s.max(-1, keepdims=True))
            pr /= pr.sum(-1, keepdims=True)
            out[:, hq] = pr @ v[:, hq // g]
        x = x + out.reshape(S, HQ * D) @ wo
        h = norm(x, n2)
        gate, up = h @ w1, h @ w3
        x = x + (gate / (1 + np.exp(-gate)) * up) @ w2
    x = norm(x, f("model.norm.weight"))
    hname = "output.weight" if arch == "internlm2" else "lm_head.weight"
    head = f(hname) if hname in t else emb
    return x @ head.T


def to_tiny(spec, cfg_tensors, tp=1, rank=0):
    """loader output -> oracle TinyWeights (TurboMind layout)."""
    t = cfg_tensors
    w = TinyWeights.__new__(TinyWeights)
    w.hidden, w.layers, w.heads, w.kv_heads, w.inter, w.vocab = H, L, HQ // tp, LD.kv_heads_for_tp(HKV, tp) // tp, I // tp, V
    T = BF16 if spec.dtype == "bf16" else F16
    w.fmt, w.dtype, w.D = spec.weight_format, T, D
    glob = LD.convert_globals(spec, t, tp, rank)
    w.embed = glob["tok_embeddings"].to(torch.float32).numpy()
    w.final_norm = glob["norm"].to(torch.float32).numpy()
    out = glob["output"].numpy()
    w.output = dict(w=out, deq=rn(out, T))
    w.L = []

    def lin(x):
        if x.quant:
            q, s, z = x.q.numpy(), x.scales.numpy(), x.zeros.numpy()
            # the TM pre-prepare packing must round-trip
            assert (LD.pack_u4_row(x.q).numpy().view(np.uint32) == OL.pack_u4_row(q).view(np.uint32)).all()
            return dict(q=q, scales=s, zeros=z, deq=OL.dequant_u4(q, s, z, F16))
        if x.fp8:
            w8, sc = x.w8.numpy(), x.scales.numpy()
            assert tuple(sc.shape) == (w8.shape[0] // 128, w8.shape[1])
            return dict(w8=w8, scales=sc, deq=OL.dequant_fp8_cols(w8, sc, 128, T))
        return dict(w=x.w.numpy(), deq=rn(x.w.numpy(), T))

    for i in range(L):
        c = LD.convert_layer(spec, t, i, tp, rank)
        extra = {"qkv_bias": rn(c["qkv_bias"].to(torch.float32).numpy(), T)} if "qkv_bias" in c else {}
        extra.update({k: rn(c[k].to(torch.float32).numpy(), T) for k in ("q_norm", "k_norm") if k in c})
        w.L.append(dict(attn_norm=c["attention_norm"].to(torch.float32).numpy(),
                        ffn_norm=c["ffn_norm"].to(torch.float32).numpy(), **extra,
                        w_qkv=lin(c["w_qkv"]), wo=lin(c["wo"]), w1w3=lin(c["w1w3"]), w2=lin(c["w2"])))
    return w


@pytest.mark.parametrize("arch", ["internlm2", "llama", "qwen2", "qwen3"])
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("tp", [1, 2, 4])        # 4 > the 2 KV heads: each is replicated (builders/attention.py:30-50)
def test_loader_transforms_preserve_model_semantics(arch, family, tp):
    cfg, t, deq = make_checkpoint(arch, family, seed=1)
    spec = LD.spec_from_config(cfg)
    assert spec.weight_format == {None: "fp16", "fp8": "fp8"}.get(family, "awq") and spec.arch == arch
    assert spec.checkpoint_format == family and spec.dtype == ("bf16" if family == "fp8" else "f16")
    assert spec.qkv_bias == (arch == "qwen2") and spec.qk_norm == (arch == "qwen3")
    T = BF16 if family == "fp8" else F16
    # logits are O(1); measured max error 5e-4 (fp16) / 5e-3 (bf16: 8 significand bits, 8x the rounding per operation)
    tokens = [5, 17, 200, 3, 99]
    ref = hf_forward(arch, t, deq, tokens)
    tol = (4e-3 if T == F16 else 3e-2) * max(1.0, float(np.abs(ref).max()))      # relative for the larger tied-head logits
    shards = [to_tiny(spec, t, tp, r) for r in range(tp)]
    full = shards[0] if tp == 1 else to_tiny(spec, t, 1, 0)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=D, base=10000.0)
    caches = [OA.PagedKvCache(2, L, OA.KvLayout(LD.kv_heads_for_tp(HKV, tp) // tp, 64, D, 16), T) for _ in range(tp)]
    for pos, tok in enumerate(tokens):
        logits, _ = decode_step(full, np.array([tok]), caches[0], [[0]], np.array([pos], np.int32), rope,
                                tp_shards=shards if tp > 1 else None, caches=caches if tp > 1 else None)
        err = np.abs(logits[0, :V] - ref[pos]).max()
        assert err < tol, (pos, err)
        assert logits[0, :V].argmax() == ref[pos].argmax() or np.sort(ref[pos])[-1] - np.sort(ref[pos])[-2] < 2 * tol


def test_rope_perm_matches_reference_definition():
    # _reorder_rotary_emb (models/utils.py:306-324): new[h, 2i+j] = old[h, j*D/2 + i]
    p = LD.rope_perm(2, 8)
    assert p.tolist() == [0, 4, 1, 5, 2, 6, 3, 7, 8, 12, 9, 13, 10, 14, 11, 15]


def test_awq_unpack_matches_oracle():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, (8, 32)).astype(np.uint8)
    w = OL.awq_pack(q)
    assert (LD.awq_unpack(torch.from_numpy(w)).numpy() == q).all()


def test_lazy_checkpoint_reads_sharded_and_indexed_directories(tmp_path):
    import json
    from safetensors.torch import save_file
    cfg, t, _ = make_checkpoint("llama", "awq", seed=3)
    names = sorted(t)
    half = len(names) // 2
    d1 = tmp_path / "headers"
    d1.mkdir()
    save_file({k: t[k].contiguous() for k in names[:half]}, str(d1 / "model-00001-of-00002.safetensors"))
    save_file({k: t[k].contiguous() for k in names[half:]}, str(d1 / "model-00002-of-00002.safetensors"))
    (d1 / "config.json").write_text(json.dumps(cfg))
    c1, lazy = LD.read_checkpoint(str(d1))
    assert c1 == cfg and len(lazy) == len(t) and set(lazy.keys()) == set(t)
    assert "nope" not in lazy and lazy.get("nope") is None
    for k in (names[0], names[-1], names[half]):
        assert torch.equal(lazy[k], t[k])
    # with an index file the weight_map decides; a shard it names but that is absent is an error up front
    d2 = tmp_path / "indexed"
    d2.mkdir()
    save_file({k: t[k].contiguous() for k in names}, str(d2 / "model-00001-of-00001.safetensors"))
    (d2 / "config.json").write_text(json.dumps(cfg))
    (d2 / "model.safetensors.index.json").write_text(json.dumps(
        {"metadata": {}, "weight_map": {k: "model-00001-of-00001.safetensors" for k in names}}))
    _, lazy2 = LD.read_checkpoint(str(d2))
    assert torch.equal(lazy2[names[1]], t[names[1]])
    (d2 / "model.safetensors.index.json").write_text(json.dumps({"weight_map": {names[0]: "gone.safetensors"}}))
    with pytest.raises(FileNotFoundError, match="gone.safetensors"):
        LD.read_checkpoint(str(d2))
    # the converted layer is the same whether it comes from the in-memory dict or the lazy reader
    spec = LD.spec_from_config(cfg)
    a, b = LD.convert_layer(spec, t, 1), LD.convert_layer(spec, lazy, 1)
    for k in ("w_qkv", "wo", "w1w3", "w2"):
        assert torch.equal(a[k].q, b[k].q) and torch.equal(a[k].scales, b[k].scales) and torch.equal(a[k].zeros, b[k].zeros)
    assert torch.equal(a["ffn_norm"], b["ffn_norm"])


def test_rope_scaling_is_parsed_like_the_reference():
    """parse_rope -> RopeSpec -> b200_rope_config against what the reference's OWN parse_rope_param + copy_rope_config
    write into its C++ RopeConfig for the same HF configs (tests/golden/ref_rope_configs.json, recorded by
    tools/make_rope_golden.py executing lmdeploy/turbomind/models/utils.py:76-187).  Includes the reference's yarn
    behaviour: the correction range uses the config's max_position_embeddings, `original_max_position_embeddings` only
    re-derives the factor (Qwen2.5-style configs end up with factor 1.0)."""
    import json
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_rope_configs.json")))
    assert len(golden) >= 14
    checked = 0
    for name, g in golden.items():
        if "raises" in g:                                          # unknown types: refused here too, loudly
            with pytest.raises(NotImplementedError):
                LD.parse_rope(g["config"], g["head_dim"]).config_fields()
            continue
        if g["rope_config"]["type"] == 3:
            # dynamic NTK: parsed field for field; the kernels are handed plain RoPE (what the reference computes for
            # prompts within the trained context), longer prompts are refused at submission
            r = LD.parse_rope(g["config"], g["head_dim"])
            want = g["rope_config"]
            assert (r.type, r.base, r.dim, r.factor, r.max_position_embeddings) == \
                ("dynamic", want["base"], want["dim"], want["factor"], want["max_position_embeddings"])
            f = r.config_fields()
            assert (f["type"], f["factor"], f["base"]) == (1, 0.0, want["base"])
            n = want["max_position_embeddings"]
            assert r.dynamic_ntk_base(n) == r.base and r.dynamic_ntk_base(1) == r.base
            # init_dynamic_ntk: base * (factor * len / max_pos - (factor - 1)) ** (dim / (dim - 2))
            assert r.dynamic_ntk_base(2 * n) == pytest.approx(r.base * (2.0 * 2 - 1.0) ** (128 / 126.0), rel=1e-12)
            checked += 5
            continue
        got = LD.parse_rope(g["config"], g["head_dim"]).config_fields()
        for field, want in g["rope_config"].items():           # every field the reference assigns for this rope type
            assert got[field] == pytest.approx(want, rel=1e-12), (name, field, got[field], want)
            checked += 1
    assert checked >= 70
    with pytest.raises(NotImplementedError):                      # multimodal rope is outside the hot-path scope
        LD.parse_rope({"max_position_embeddings": 4096, "rope_scaling": {"type": "mrope", "mrope_section": [16, 24, 24]}}, 128)


def test_partial_rotary_reorders_only_the_rotated_dims():
    p = LD.rope_perm(1, 8, 4).tolist()
    assert p == [0, 2, 1, 3, 4, 5, 6, 7]


def test_architecture_admission():
    base = {"hidden_size": 512, "num_hidden_layers": 2, "num_attention_heads": 4, "num_key_value_heads": 2,
            "intermediate_size": 512, "vocab_size": 256, "max_position_embeddings": 32768}
    assert LD.spec_from_config(dict(base, architectures=["MistralForCausalLM"], sliding_window=None)).arch == "llama"
    assert LD.spec_from_config(dict(base, architectures=["Qwen2ForCausalLM"])).qkv_bias
    assert LD.spec_from_config(dict(base, architectures=["LlamaForCausalLM"], attention_bias=True)).qkv_bias
    for bad in (dict(architectures=["MistralForCausalLM"], sliding_window=4096),
                dict(architectures=["Qwen2ForCausalLM"], use_sliding_window=True),
                dict(architectures=["MixtralForCausalLM"]), dict(architectures=["Qwen2MoeForCausalLM"]),
                dict(architectures=["LlamaForCausalLM"], mlp_bias=True)):
        with pytest.raises(NotImplementedError):
            LD.spec_from_config(dict(base, **bad))


def test_kv_heads_are_replicated_to_a_tp_multiple():
    assert [LD.kv_heads_for_tp(8, tp) for tp in (1, 2, 4, 8)] == [8, 8, 8, 8]
    assert LD.kv_heads_for_tp(4, 8) == 8 and LD.kv_heads_for_tp(2, 8) == 8 and LD.kv_heads_for_tp(1, 4) == 4
    with pytest.raises(NotImplementedError):
        LD.kv_heads_for_tp(3, 2)                 # 4 is not a multiple of 3: the reference asserts here too
    # the replicated layout: rank r of tp=4 over 2 KV heads reads KV head r // 2 -- the one its q head belongs to
    cfg, t, _ = make_checkpoint("llama", None, seed=3)
    spec = LD.spec_from_config(cfg)
    full = LD.convert_layer(spec, t, 0, 1, 0)["w_qkv"].w
    for r in range(4):
        shard = LD.convert_layer(spec, t, 0, 4, r)["w_qkv"].w
        assert tuple(shard.shape) == (H, 3 * D)
        assert torch.equal(shard[:, :D], full[:, r * D:(r + 1) * D])                                    # q head r
        assert torch.equal(shard[:, D:2 * D], full[:, (HQ + r // 2) * D:(HQ + r // 2 + 1) * D])          # its K head
        assert torch.equal(shard[:, 2 * D:], full[:, (HQ + HKV + r // 2) * D:(HQ + HKV + r // 2 + 1) * D])


def test_column_layouts_match_the_reference_loader_functions():
    """w_qkv / w1w3 column order per TP rank against the reference's OWN tensor functions (_split_qkv_gqa,
    _reorder_rotary_emb, _repeat_kv_heads, fuse_qkv, _interleave_w1w3) executed on index tensors by
    tools/make_layout_golden.py -> tests/golden/ref_loader_layouts.json.  Here convert_layer runs on checkpoints whose
    weights ARE the source column ids."""
    import json
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_loader_layouts.json")))
    hidden, inter = 128, 512
    checked = 0

    def checkpoint(arch, Hq, Hkv, Dh):
        ids = lambda lo, n: (lo + torch.arange(n, dtype=torch.float32))[:, None].repeat(1, hidden)      # [out, in]
        o = torch.zeros(hidden, Hq * Dh)
        gate, up, down = ids(0, inter), ids(inter, inter), torch.zeros(hidden, inter)
        ones = torch.ones(hidden)
        p = "model.layers.0."
        if arch == "internlm2":
            return {p + "attention.wqkv.weight": ids(0, (Hq + 2 * Hkv) * Dh), p + "attention.wo.weight": o,
                    p + "feed_forward.w1.weight": gate, p + "feed_forward.w3.weight": up, p + "feed_forward.w2.weight": down,
                    p + "attention_norm.weight": ones, p + "ffn_norm.weight": ones}
        return {p + "self_attn.q_proj.weight": ids(0, Hq * Dh), p + "self_attn.k_proj.weight": ids(Hq * Dh, Hkv * Dh),
                p + "self_attn.v_proj.weight": ids((Hq + Hkv) * Dh, Hkv * Dh), p + "self_attn.o_proj.weight": o,
                p + "mlp.gate_proj.weight": gate, p + "mlp.up_proj.weight": up, p + "mlp.down_proj.weight": down,
                p + "input_layernorm.weight": ones, p + "post_attention_layernorm.weight": ones}

    def spec_for(arch, Hq, Hkv, Dh, rope_dim):
        return LD.spec_from_config({"architectures": ["InternLM2ForCausalLM" if arch == "internlm2" else "LlamaForCausalLM"],
                                    "hidden_size": hidden, "num_hidden_layers": 1, "num_attention_heads": Hq,
                                    "num_key_value_heads": Hkv, "head_dim": Dh, "intermediate_size": inter, "vocab_size": 128,
                                    "partial_rotary_factor": rope_dim / Dh, "torch_dtype": "float32"})

    for g in golden["qkv"]:
        spec = spec_for(g["arch"], g["heads"], g["kv_heads"], g["head_dim"], g["rope_dim"])
        t = checkpoint(g["arch"], g["heads"], g["kv_heads"], g["head_dim"])
        for r, want in enumerate(g["ranks"]):
            w = LD.convert_layer(spec, t, 0, g["tp"], r)["w_qkv"].w
            assert (w == w[0]).all()
            assert w[0].to(torch.int64).tolist() == want, (g["arch"], g["heads"], g["kv_heads"], g["rope_dim"], g["tp"], r)
            checked += 1
    for g in golden["ffn"]:
        spec, t = spec_for("llama", 4, 2, 128, 128), checkpoint("llama", 4, 2, 128)
        for r, want in enumerate(g["ranks"]):
            assert g["inter"] == inter
            w = LD.convert_layer(spec, t, 0, g["tp"], r)["w1w3"].w
            assert w[0].to(torch.int64).tolist() == want, ("ffn", g["tp"], r)
            checked += 1
    assert checked >= 30


def test_engine_defaults_follow_the_reference():
    """session_len / max_batch_size when the user leaves them unset: tests/golden/ref_engine_defaults.json holds what the
    reference's own _get_and_verify_max_len and get_max_batch_size return (tools/make_engine_defaults_golden.py)."""
    import json
    from lmdeploy_b200.messages import DEFAULT_MAX_BATCH_SIZE
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_engine_defaults.json")))
    assert len(g["session_len"]) >= 7
    for name, case in g["session_len"].items():
        assert LD.default_session_len(case["config"]) == case["value"], name
    assert DEFAULT_MAX_BATCH_SIZE == g["max_batch_size"]["NVIDIA B200"]
    cfg, _, _ = make_checkpoint("llama", None, seed=0)
    cfg = dict(cfg, max_position_embeddings=131072, model_max_length=8192)
    spec = LD.spec_from_config(cfg)
    assert spec.session_len == 8192 and spec.max_position_embeddings == 131072       # RoPE still sees the config's own value


def test_hf_overrides_are_merged_into_the_config():
    cfg = {"a": 1, "rope_scaling": {"type": "linear", "factor": 2.0}, "text_config": {"x": {"y": 1, "z": 2}}}
    out = LD.deep_merge(cfg, {"rope_scaling": {"type": "dynamic"}, "text_config": {"x": {"y": 5}}, "new": [1]})
    assert out is cfg and cfg == {"a": 1, "rope_scaling": {"type": "dynamic", "factor": 2.0},
                                  "text_config": {"x": {"y": 5, "z": 2}}, "new": [1]}


def test_biases_follow_the_checkpoint_and_internlm3_is_a_llama():
    cfg, t, _ = make_checkpoint("qwen2", None, seed=2)
    spec = LD.spec_from_config(dict(cfg, architectures=["LlamaForCausalLM"]))          # config says nothing about biases
    assert not spec.qkv_bias and LD.with_checkpoint_biases(spec, t).qkv_bias            # the tensors do
    cfg, t, _ = make_checkpoint("llama", None, seed=2)
    spec = LD.spec_from_config(dict(cfg, attention_bias=True))
    assert spec.qkv_bias and not LD.with_checkpoint_biases(spec, t).qkv_bias
    spec = LD.spec_from_config(dict(cfg, architectures=["InternLM3ForCausalLM"],
                                    rope_scaling={"rope_type": "dynamic", "factor": 6.0}, max_position_embeddings=32768))
    assert spec.arch == "llama" and spec.rope.type == "dynamic" and spec.session_len == 32768


def test_oracle_qk_rmsnorm_is_a_per_head_rmsnorm():
    """The grouped-sum restatement (the compiled reference kernel's order) against a plain float64 RMSNorm per head."""
    rng = np.random.default_rng(5)
    Hq, Hkv = 4, 2
    for T in (F16, BF16):
        qkv = rn(rng.standard_normal((7, (Hq + 2 * Hkv) * 128)).astype(np.float32) * 3, T)
        qw = rn((1 + 0.2 * rng.standard_normal(128)).astype(np.float32), T)
        kw = rn((1 + 0.2 * rng.standard_normal(128)).astype(np.float32), T)
        got = OL.qk_rmsnorm(qkv, qw, kw, Hq, Hkv, 128, 1e-6, T)
        x = qkv[:, :(Hq + Hkv) * 128].reshape(7, Hq + Hkv, 128).astype(np.float64)
        w = np.concatenate([np.tile(qw, (Hq, 1)), np.tile(kw, (Hkv, 1))]).astype(np.float64)
        want = x / np.sqrt((x ** 2).mean(-1, keepdims=True) + 1e-6) * w
        err = np.abs(got[:, :(Hq + Hkv) * 128].reshape(7, Hq + Hkv, 128) - want)
        assert err.max() <= (4e-3 if T == F16 else 3e-2)            # two roundings to T of O(1..4) values
        assert np.array_equal(got[:, (Hq + Hkv) * 128:], qkv[:, (Hq + Hkv) * 128:])      # v untouched
        assert np.array_equal(rn(got, T), got)                                          # exact in T
