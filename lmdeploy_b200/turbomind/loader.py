"""HF checkpoint -> TurboMind pre-prepare tensors of one TP rank (then b200_linear_prepare on the GPU).

Restates the loader transforms of the reference for the two architectures on the hot path
(lmdeploy/turbomind/models/internlm2.py:32-135, models/llama.py:44-99):
  * every linear is read through checkpoint_formats.FormatResolver: AWQ 'gemm', GPTQ (sym, no act-order) and
    compressed-tensors pack-quantized int4 all become uint4 [K,N] + scales/zeros [K/128,N] and are re-packed
    row-plain (weight_format.py:47-74, 200-346); fp16 HF linears [N,K] are transposed to [K,N];
  * InternLM2 `wqkv` is de-interleaved from the per-kv-group layout [q0..q(g-1) k v] (internlm2.py:32-49);
  * q / k output channels are permuted from HF half-split RoPE to TurboMind's adjacent-pair RoPE
    (models/utils.py:306-324) -- for quantised weights the same column permutation is applied to qweight,
    scales and zeros (utils.py:355-371);
  * gate/up are element-interleaved for the fused SiLU epilogue (builders/ffn.py:31-34);
  * TP: q/k/v heads, gate/up columns, wo/w2 rows and lm_head vocab are sliced per rank
    (builders/_base.py:102-113, attention.py:65-76, ffn.py:238-243).
All functions are pure tensor code (CPU-testable); only `load_hf_checkpoint` touches the GPU."""
import json
import math
import os

import torch

from .checkpoint_formats import (AWQ_NIBBLE_OF_COLUMN, INT4_METHODS, FormatResolver, Lin, pack_u4_row,  # noqa: F401
                                 parse_quantization_config, unpack_last)
from .specs import ModelSpec, RopeSpec

_Lin = Lin


def awq_unpack(x):
    """int32 [R, C/8] -> uint8 [R, C]   (_unpack_awq_gemm, weight_format.py:56-60)."""
    return unpack_last(x, AWQ_NIBBLE_OF_COLUMN)


def rope_perm(num_heads, head_dim, rope_dim=None):
    """Column permutation of _reorder_rotary_emb: new[h, 2i+j] = old[h, j*D/2 + i]."""
    rope_dim = rope_dim or head_dim
    idx = torch.arange(num_heads * head_dim).view(num_heads, head_dim)
    rot = idx[:, :rope_dim].reshape(num_heads, 2, rope_dim // 2).transpose(1, 2).reshape(num_heads, rope_dim)
    return torch.cat([rot, idx[:, rope_dim:]], dim=1).reshape(-1)


def resolver_for(spec: ModelSpec):
    return FormatResolver(None if spec.weight_format == "fp16" else (spec.checkpoint_format or spec.weight_format))


def _yarn_mscale(scale, mscale=1.0):
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


def parse_rope(cfg, head_dim):
    """HF config -> RopeSpec, field for field what parse_rope_param produces (lmdeploy/turbomind/models/utils.py:76-166):
    transformers >= 5 `rope_parameters` or the older `rope_theta` + `rope_scaling`; partial_rotary_factor; linear / llama3 /
    yarn scaling (yarn: attention_factor from mscale / mscale_all_dim when not given, factor re-derived from the original
    context length)."""
    rp = cfg.get("rope_parameters")
    if rp is not None:
        scaling, theta = rp, float(rp.get("rope_theta", 10000.0))
    else:
        scaling, theta = cfg.get("rope_scaling"), float(cfg.get("rope_theta", 10000.0))
    max_pos = int(cfg.get("max_position_embeddings", 0))
    prf = (rp or {}).get("partial_rotary_factor")
    if prf is None:
        prf = float(cfg.get("partial_rotary_factor", 1.0))
    # `factor` 1.0 and the MODEL's context length for every type: what copy_rope_config hands the C++ config
    # (models/utils.py:170-187).  Note the yarn quirk this preserves: the correction range is computed from the
    # config's max_position_embeddings, not from rope_scaling.original_max_position_embeddings (which only re-derives
    # the factor) -- tests/golden/ref_rope_configs.json records the reference's own outputs.
    kw = dict(type="default", dim=int(head_dim * prf), base=theta, factor=1.0, max_position_embeddings=max_pos)
    if scaling:
        rtype = scaling.get("rope_type") or scaling.get("type") or "default"
        factor = float(scaling.get("factor", 0.0) or 0.0)
        if rtype in ("default", "mrope"):
            if scaling.get("mrope_section"):
                raise NotImplementedError("multimodal rope (mrope_section) is outside the hot-path scope")
        elif rtype == "linear":
            kw.update(type="linear", factor=factor)
        elif rtype == "llama3":
            kw.update(type="llama3", factor=factor, low_freq_factor=float(scaling.get("low_freq_factor", 1.0)),
                      high_freq_factor=float(scaling.get("high_freq_factor", 1.0)),
                      original_max_position_embeddings=int(scaling.get("original_max_position_embeddings", 0)))
        elif rtype == "yarn":
            af = scaling.get("attention_factor")
            if af is None:
                ms, msa = scaling.get("mscale"), scaling.get("mscale_all_dim")
                af = _yarn_mscale(factor, ms) / _yarn_mscale(factor, msa) if ms is not None and msa is not None \
                    else _yarn_mscale(factor)
            orig = scaling.get("original_max_position_embeddings")
            if orig is not None:
                factor = max_pos / orig
            kw.update(type="yarn", factor=float(factor), attention_factor=float(af),
                      beta_fast=float(scaling.get("beta_fast", 32.0)), beta_slow=float(scaling.get("beta_slow", 1.0)))
        elif rtype == "dynamic":
            # dynamic NTK: a per-REQUEST base, fixed at admission from the prompt length (init_dynamic_ntk,
            # unified_attention_layer.cc:228-243); prompts within the trained context keep the config's base, which is
            # the only case served here (engine.TurboMindInstance refuses longer prompts, it does not mis-rotate them)
            kw.update(type="dynamic", factor=factor)
        else:
            raise NotImplementedError(f"unsupported rope type {rtype!r}")
    return RopeSpec(**kw)


_CONTEXT_LENGTH_KEYS = ("max_position_embeddings", "n_positions", "max_seq_len", "seq_length", "model_max_length",
                        "max_sequence_length", "max_seq_length", "seq_len")


def default_session_len(cfg):
    """Session length when the user sets none: the minimum over the context-length keys the config carries (looked up
    in the language-model sub-config when there is one), 2048 when it carries none -- lmdeploy/utils.py:321-391
    `_get_and_verify_max_len(hf_config, None)`, pinned by tests/golden/ref_engine_defaults.json."""
    for key in ("language_config", "llm_config", "text_config"):
        cfg = cfg.get(key, cfg)
    found = [cfg[k] for k in _CONTEXT_LENGTH_KEYS if cfg.get(k) is not None]
    return int(min(found)) if found else 2048


def spec_from_config(cfg):
    arch = cfg["architectures"][0]
    # lmdeploy/turbomind/supported_models.py SUPPORTED_ARCHS (dense text models of the hot-path scope); Mistral dense uses
    # the Llama tensor names (no sliding window, below)
    name = {"InternLM2ForCausalLM": "internlm2", "LlamaForCausalLM": "llama", "InternLM3ForCausalLM": "llama",
            "Qwen2ForCausalLM": "qwen2", "Qwen3ForCausalLM": "qwen3", "MistralForCausalLM": "llama"}.get(arch)
    if name is None:
        raise NotImplementedError(f"architecture {arch} is outside the hot-path scope (InternLM2 / Llama / Qwen2 / Qwen3 dense)")
    window = cfg.get("sliding_window")
    if cfg.get("num_experts") or cfg.get("num_local_experts") or cfg.get("use_sliding_window") or \
            (arch == "MistralForCausalLM" and window and window < cfg.get("max_position_embeddings", 0)):
        raise NotImplementedError("MoE / sliding-window variants are outside the hot-path scope")
    # q/k/v bias: always for Qwen2 (lmdeploy/turbomind/models/qwen2.py:78-97 bundles it with the linears), optional for
    # Llama-style configs; a bias on o_proj / mlp is not supported
    qkv_bias = name == "qwen2" or bool(cfg.get("attention_bias", False))
    if cfg.get("mlp_bias", False):
        raise NotImplementedError("mlp_bias checkpoints are not supported")
    method, _group = parse_quantization_config(cfg.get("quantization_config"))
    # one on-GPU format per weight family: every int4 checkpoint becomes 'awq' (u4 g128), fp8 stays fp8
    fmt = "awq" if method in INT4_METHODS else "fp8" if method == "fp8" else "fp16"
    # compute dtype: int4 formats force fp16 (converter.py:38-48); otherwise the checkpoint's own dtype
    hf_dtype = str(cfg.get("dtype", cfg.get("torch_dtype", "float16"))).replace("torch.", "")
    dtype = "f16" if fmt == "awq" or hf_dtype != "bfloat16" else "bf16"
    head_dim = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
    rope = parse_rope(cfg, head_dim)
    return ModelSpec(name=os.path.basename(cfg.get("_name_or_path", name)) or name, arch=name,
                     hidden=cfg["hidden_size"], layers=cfg["num_hidden_layers"], heads=cfg["num_attention_heads"],
                     kv_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]),
                     inter=cfg["intermediate_size"], vocab=cfg["vocab_size"],
                     rope_theta=rope.base, head_dim=head_dim,
                     norm_eps=float(cfg.get("rms_norm_eps", 1e-5)), rope_type=rope.type, rope_factor=rope.factor, rope=rope,
                     weight_format=fmt,
                     checkpoint_format=method, dtype=dtype, qkv_bias=qkv_bias,
                     max_position_embeddings=int(cfg.get("max_position_embeddings", 0)),
                     session_len=default_session_len(cfg), qk_norm=name == "qwen3")


def kv_heads_for_tp(kv_heads, tp):
    """KV head count after the replication that makes it divisible by the TP size (the reference's loader repeats each
    KV head, lmdeploy/turbomind/builders/attention.py:30-50 `_repeat_kv_heads`: 4 KV heads on 8 GPUs become 8)."""
    if kv_heads % tp == 0:
        return kv_heads
    target = (kv_heads + tp - 1) // tp * tp
    if target % kv_heads:
        raise NotImplementedError(f"{kv_heads} KV heads cannot be replicated evenly over tp={tp}")
    return target


def convert_layer(spec: ModelSpec, tensors, i, tp=1, rank=0):
    """-> dict(w_qkv, wo, w1w3, w2 : _Lin ; attention_norm, ffn_norm : tensor) for TP rank `rank`."""
    D, Hq, Hkv = spec.head_dim, spec.heads, spec.kv_heads
    rd = resolver_for(spec)

    def _get_linear(prefix):
        return rd.read(tensors, prefix)

    if spec.arch == "internlm2":
        p = f"model.layers.{i}."
        wqkv = _get_linear(p + "attention.wqkv")
        g = Hq // Hkv
        idx = torch.arange(Hkv * (g + 2) * D).view(Hkv, g + 2, D)
        q = wqkv.cols(idx[:, :g].reshape(-1))
        k = wqkv.cols(idx[:, g].reshape(-1))
        v = wqkv.cols(idx[:, g + 1].reshape(-1))
        wo = _get_linear(p + "attention.wo")
        w1 = _get_linear(p + "feed_forward.w1")
        w3 = _get_linear(p + "feed_forward.w3")
        w2 = _get_linear(p + "feed_forward.w2")
        an, fn = tensors[p + "attention_norm.weight"], tensors[p + "ffn_norm.weight"]
    else:
        p = f"model.layers.{i}."
        q = _get_linear(p + "self_attn.q_proj")
        k = _get_linear(p + "self_attn.k_proj")
        v = _get_linear(p + "self_attn.v_proj")
        wo = _get_linear(p + "self_attn.o_proj")
        w1 = _get_linear(p + "mlp.gate_proj")
        w3 = _get_linear(p + "mlp.up_proj")
        w2 = _get_linear(p + "mlp.down_proj")
        an, fn = tensors[p + "input_layernorm.weight"], tensors[p + "post_attention_layernorm.weight"]
    rope_dim = spec.rope.dim if spec.rope is not None else D        # partial rotary: only the rotated dims are reordered
    q = q.cols(rope_perm(Hq, D, rope_dim))
    k = k.cols(rope_perm(Hkv, D, rope_dim))
    rep = kv_heads_for_tp(Hkv, tp) // Hkv
    if rep > 1:     # every KV head `rep` times in a row: q head j keeps reading (a copy of) kv head j // (Hq / Hkv)
        idx = torch.arange(Hkv * D).view(Hkv, D).repeat_interleave(rep, 0).reshape(-1)
        k, v = k.cols(idx), v.cols(idx)
    hq, hkv, it = Hq // tp, Hkv * rep // tp, spec.inter // tp
    qs = q.cols(torch.arange(rank * hq * D, (rank + 1) * hq * D))
    ks = k.cols(torch.arange(rank * hkv * D, (rank + 1) * hkv * D))
    vs = v.cols(torch.arange(rank * hkv * D, (rank + 1) * hkv * D))
    cols = torch.arange(rank * it, (rank + 1) * it)
    gate, up = w1.cols(cols), w3.cols(cols)
    inter_idx = torch.stack([torch.arange(it), it + torch.arange(it)], 1).reshape(-1)
    w_qkv = _Lin.cat([qs, ks, vs])
    if spec.qkv_bias and w_qkv.bias is None:
        raise KeyError(f"layer {i}: the config announces q/k/v biases but the checkpoint has none")
    if wo.bias is not None or w2.bias is not None or gate.bias is not None:
        raise NotImplementedError(f"layer {i}: biases on o_proj / mlp linears are not supported")
    out = dict(w_qkv=w_qkv, wo=wo.rows(rank * hq * D, (rank + 1) * hq * D),
               w1w3=_Lin.cat([gate, up]).cols(inter_idx), w2=w2.rows(rank * it, (rank + 1) * it),
               attention_norm=an, ffn_norm=fn)
    if spec.qkv_bias:
        out["qkv_bias"] = w_qkv.bias          # applied by the attention prologue, not by the GEMM
        w_qkv.bias = None
    if spec.qk_norm:
        # [head_dim] weights shared by every head; their elements follow the RoPE reorder of the channels they scale
        # (lmdeploy/turbomind/models/qwen3.py reorders q_norm / k_norm with reorder_rotary_emb like the projections)
        perm = rope_perm(1, D, rope_dim)
        out["q_norm"] = tensors[p + "self_attn.q_norm.weight"][perm]
        out["k_norm"] = tensors[p + "self_attn.k_norm.weight"][perm]
    return out


def convert_globals(spec: ModelSpec, tensors, tp=1, rank=0):
    if spec.arch == "internlm2":
        emb, norm, out = tensors["model.tok_embeddings.weight"], tensors["model.norm.weight"], tensors["output.weight"]
    else:
        emb, norm = tensors["model.embed_tokens.weight"], tensors["model.norm.weight"]
        out = tensors.get("lm_head.weight", emb)
    vpad = (spec.vocab + 128 * tp - 1) // (128 * tp) * (128 * tp)
    vl = vpad // tp
    o = torch.zeros(vpad, spec.hidden, dtype=torch.float32)
    o[:spec.vocab] = out.to(torch.float32)
    return dict(tok_embeddings=emb, norm=norm, output=o[rank * vl:(rank + 1) * vl].t().contiguous())


class LazyCheckpoint:
    """Read-only name -> tensor mapping over the *.safetensors shards of a checkpoint directory.  Tensors are read from
    disk on access (safetensors memory-maps the shard) and not retained, so converting layer i holds only layer i in
    host memory -- a 70B checkpoint does not fit twice.  Shard lookup follows `model.safetensors.index.json` when it
    exists (its `weight_map`), else the headers of every shard (lmdeploy/turbomind/checkpoint.py does the same
    per-shard iteration through safe_open)."""

    def __init__(self, path):
        from safetensors import safe_open
        self._open = safe_open
        self.path = path
        self._shard_of = {}
        index = os.path.join(path, "model.safetensors.index.json")
        shards = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
        if not shards:
            raise FileNotFoundError(f"no *.safetensors files in {path}")
        if os.path.exists(index):
            with open(index) as fh:
                self._shard_of = dict(json.load(fh)["weight_map"])
            missing = sorted(set(self._shard_of.values()) - set(shards))
            if missing:
                raise FileNotFoundError(f"{index} names shards that are not in {path}: {missing}")
        else:
            for f in shards:
                with safe_open(os.path.join(path, f), framework="pt", device="cpu") as fh:
                    for k in fh.keys():
                        self._shard_of[k] = f
        self._handles = {}

    def _handle(self, shard):
        h = self._handles.get(shard)
        if h is None:
            h = self._handles[shard] = self._open(os.path.join(self.path, shard), framework="pt", device="cpu")
        return h

    def __contains__(self, name):
        return name in self._shard_of

    def __getitem__(self, name):
        return self._handle(self._shard_of[name]).get_tensor(name)

    def get(self, name, default=None):
        return self[name] if name in self._shard_of else default

    def keys(self):
        return self._shard_of.keys()

    def __len__(self):
        return len(self._shard_of)


def read_checkpoint(path):
    """-> (config dict, LazyCheckpoint)."""
    with open(os.path.join(path, "config.json")) as fh:
        cfg = json.load(fh)
    return cfg, LazyCheckpoint(path)


def with_checkpoint_biases(spec, tensors):
    """q/k/v biases are a property of the CHECKPOINT: the reference bundles whatever `.bias` tensors exist with their
    linears (lmdeploy/turbomind/models/llama.py reads them through `_linear(prefix)`), whatever the config says."""
    import dataclasses
    key = "model.layers.0.attention.wqkv.bias" if spec.arch == "internlm2" else "model.layers.0.self_attn.q_proj.bias"
    has = key in tensors
    return spec if has == spec.qkv_bias else dataclasses.replace(spec, qkv_bias=has)


def deep_merge(base, override):
    """TurbomindEngineConfig.hf_overrides: merged into config.json key by key, dicts recursively
    (lmdeploy/turbomind/converter.py:59-70 `_deep_merge`), e.g. {"rope_scaling": {"type": "dynamic", "factor": 2.0}}."""
    for k, v in override.items():
        if isinstance(base.get(k), dict) and isinstance(v, dict):
            deep_merge(base[k], v)
        else:
            base[k] = v
    return base


def load_hf_checkpoint(path, *, quant_policy=0, block_len=64, max_batch=64, tp=1, rank=0, hf_overrides=None, model_format=None):
    from .runtime import DecodeModel, _TORCH_T
    cfg, tensors = read_checkpoint(path)
    if hf_overrides:
        deep_merge(cfg, hf_overrides)
    method = (cfg.get("quantization_config") or {}).get("quant_method")
    if method is not None and model_format not in (None, method):
        # get_tm_config: the user's model_format must agree with the checkpoint's own quantization_config (converter.py)
        raise ValueError(f'mismatched quant method: user input "{model_format}" vs model quant_config "{method}"')
    if method is None and model_format not in (None, "hf"):
        raise ValueError(f'model_format "{model_format}" but the checkpoint carries no quantization_config')
    spec = spec_from_config(cfg)
    spec = with_checkpoint_biases(spec, tensors)
    import dataclasses
    # the device model sees the replicated KV head count; convert_layer replicates from the checkpoint's own
    model = DecodeModel(dataclasses.replace(spec, kv_heads=kv_heads_for_tp(spec.kv_heads, tp)), quant_policy=quant_policy,
                        block_len=block_len, max_batch=max_batch, tp_size=tp, tp_rank=rank)
    dev, T = model.device, _TORCH_T[spec.dtype]
    glob = convert_globals(spec, tensors, tp, rank)
    cache = {}

    def provider(name):
        if name in ("tok_embeddings", "norm"):
            return glob[name]
        if name == "output":
            return dict(weight=glob["output"].to(dev, T).contiguous())
        parts = name.split(".")
        i, kind = int(parts[1]), parts[2]
        if i not in cache:
            cache.clear()
            cache[i] = convert_layer(spec, tensors, i, tp, rank)
        x = cache[i][kind]
        return x.to_tm(dev, T) if isinstance(x, _Lin) else x            # norms / qkv_bias: plain tensors

    model.load_tm_weights(provider)
    eos = cfg.get("eos_token_id")
    if isinstance(eos, list):
        eos = eos[0]
    return spec, model, eos
