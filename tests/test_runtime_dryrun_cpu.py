"""Dry run of the GPU-only HOST code on the CPU: DecodeModel (weight upload through the C ABI), StaticBatchRunner (the
bench loop) and the TurboMind engine's device step, with the C library replaced by a recorder and torch.cuda by stand-ins.

No arithmetic is checked here (the GPU tests do that); what this catches before a GPU box is spent on it is the class of
mistake the host code can make on its own: a misspelt keyword into a ctypes structure, a tensor of the wrong shape or
dtype handed to `b200_model_set_*`, a call with the wrong number of arguments (the recorder checks every call against the
argtypes the binding declares), a buffer that is too small for the step it is given."""
import types

import numpy as np
import pytest
import torch

from lmdeploy_b200 import _lib as L


from dryrun import install  # noqa: E402


@pytest.fixture
def dry(monkeypatch, lib):
    """`lib` (conftest) is the real library: loaded on the CPU for its symbol table and argtypes only."""
    return install(monkeypatch.setattr, lib)


def tiny(fmt, **kw):
    from lmdeploy_b200.turbomind.specs import ModelSpec
    return ModelSpec("tiny", "llama", 512, 2, 4, 2, 1024, 1000, 1e4, weight_format=fmt, **kw)


@pytest.mark.parametrize("fmt,kw", [("awq", {}), ("fp16", {}), ("fp8", dict(dtype="bf16")), ("awq", dict(qkv_bias=True)),
                                    ("awq", dict(qk_norm=True))])
def test_model_upload_and_step_calls(dry, fmt, kw):
    from lmdeploy_b200.turbomind.runtime import DecodeModel
    spec = tiny(fmt, **kw)
    m = DecodeModel(spec, quant_policy=8, max_batch=8, device="cpu")
    m.load_synthetic(seed=0)
    names = dry.names()
    per_layer = 4
    assert names.count("b200_model_set_linear") == spec.layers * per_layer + 1            # + lm_head
    assert names.count("b200_linear_prepare") == spec.layers * per_layer + 1
    assert names.count("b200_model_set_tensor") == 2 + spec.layers * (2 + (1 if spec.qkv_bias else 0) + (2 if spec.qk_norm else 0))
    if spec.qk_norm:
        ids = [a[2] for n, a in dry.calls if n == "b200_model_set_tensor"]
        assert ids.count(L.T_Q_NORM) == ids.count(L.T_K_NORM) == spec.layers
    fmts = [a[0]._obj.weight_format for n, a in dry.calls if n == "b200_linear_prepare"]
    want = {"awq": L.FMT_U4, "fp16": L.FMT_DENSE, "fp8": L.FMT_FP8_COL}[fmt]
    assert fmts[:-1] == [want] * (spec.layers * per_layer) and fmts[-1] == L.FMT_DENSE    # lm_head is dense T
    # vocabulary padded to the GEMM's 128-column tile
    lm = [a for n, a in dry.calls if n == "b200_linear_prepare"][-1][0]._obj
    assert (lm.input_dim, lm.output_dim) == (512, 1024)
    # the rope block of a spec without an explicit RopeSpec: the literal the GPU-verified runs were made with
    r = [a[1]._obj for n, a in dry.calls if n == "b200_model_create"][-1].rope
    assert (r.type, r.dim, r.base, r.factor, r.max_position_embeddings, r.llama3_low_freq_factor, r.llama3_high_freq_factor,
            r.llama3_original_max_position_embeddings, r.yarn_attention_factor, r.yarn_beta_fast, r.yarn_beta_slow,
            r.exact_math) == (1, 128, 1e4, 0.0, 0, 1.0, 4.0, 8192, 1.0, 32.0, 1.0, 0)
    B = 3
    i32 = lambda n: torch.zeros(n, dtype=torch.int32)
    m.decode_step(batch=B, token_ids=i32(B), cu_q_len=i32(B + 1), cu_k_len=i32(B + 1), block_ptrs=torch.zeros(6, dtype=torch.int64),
                  cu_block_nums=i32(B + 1), out_tokens=i32(B), max_k_len=70)
    name, args = dry.calls[-1]
    si = args[1]._obj
    assert name == "b200_model_decode_step" and si.batch_size == B and si.max_k_len == 70 and si.advance == 0
    assert si.layer_stride_bytes == m.layer_bytes and si.finished is None
    m.decode_step(batch=B, token_ids=i32(7), cu_q_len=i32(B + 1), cu_k_len=i32(B + 1), block_ptrs=torch.zeros(6, dtype=torch.int64),
                  cu_block_nums=i32(B + 1), out_tokens=i32(7), max_k_len=70, token_num=7)
    assert dry.calls[-1][0] == "b200_model_chunk_step" and dry.calls[-1][1][2] == 7


def test_static_batch_runner_bench_loop(dry):
    """The calls bench.py makes, in order, on the device-resident and the host-fed path."""
    from lmdeploy_b200.engine import StaticBatchRunner
    from lmdeploy_b200.turbomind.runtime import DecodeModel
    m = DecodeModel(tiny("awq"), quant_policy=8, max_batch=4, device="cpu")
    m.load_synthetic(seed=0)
    r = StaticBatchRunner(m, batch=4, max_len=200, seed=0)
    assert r.block_ptrs.numel() == 4 * 4 and len(set(r.block_ptrs.tolist())) == 16
    r.prefill_random(100)
    assert r.cu_k.tolist() == [0, 101, 202, 303, 404] and r.cur_len == 101
    n0 = len(dry.calls)
    step = r.make_device_step(graph=True)
    for _ in range(3):
        step()
    names = dry.names()[n0:]
    assert names.count("b200_model_decode_step") == 2 and names.count("graph.replay") == 3       # warm-up + capture, then replays
    assert r.cur_len == 101 + 1 + 3
    out = r.step_host(np.array([1, 2, 3, 4], np.int32))
    assert out.shape == (4,) and dry.names()[-1] == "graph.replay"
    # bench.py's roofline block: the main kernel alone, then the whole operator WITH the step's plan (as a decode step calls it)
    class FakeEvent:
        def __init__(self, enable_timing=False):
            pass

        def record(self, *a):
            pass

        def elapsed_time(self, other):
            return 1.0
    import pytest
    mp = pytest.MonkeyPatch()
    mp.setattr(torch.cuda, "Event", FakeEvent)
    try:
        n1 = len(dry.calls)
        t = r.time_attention_kernel(150, iters=2)
        assert t["bytes"] > 0 and "b200_decode_attention_plan" not in dry.names()[n1:]
        calls = [c for c in dry.calls[n1:] if c[0] == "b200_decode_attention"]      # the recorded struct is the live one
        assert len(calls) == 1 + 3 + 2 and calls[-1][1][0]._obj.debug_flags == 7 and not calls[-1][1][0]._obj.plan
        n2 = len(dry.calls)
        t = r.time_attention_kernel(150, iters=2, full=True)
        names = dry.names()[n2:]
        assert names.count("b200_decode_attention_plan") == 1 and names.count("b200_decode_attention") == 1 + 3 + 2
        assert names.index("b200_decode_attention_plan") < names.index("b200_decode_attention")
        calls = [c for c in dry.calls[n2:] if c[0] == "b200_decode_attention"]
        assert calls[-1][1][0]._obj.debug_flags == 0 and calls[-1][1][0]._obj.plan
        assert r._attn_params[5] is not None
    finally:
        mp.undo()
    sb = m.step_bytes(4, 150)
    assert sb["total"] == sum(v for k, v in sb.items() if k != "total") and sb["kv_read"] > 0
    assert r.launches_per_step == 4 + 9 * 2 + 4


def test_engine_device_step_marshals_what_the_scheduler_formed(dry, monkeypatch):
    """TurboMind._execute: the arrays of a step land in the device buffers the C call receives, pure decode steps replay a
    graph from their second occurrence, a mixed step goes through the chunk entry."""
    from lmdeploy_b200.messages import GenerationConfig, TurbomindEngineConfig
    from lmdeploy_b200.turbomind.engine import TurboMind, TurboMindInstance
    monkeypatch.setenv("B200_CHUNKED_PREFILL", "4")
    from lmdeploy_b200.turbomind.runtime import DecodeModel
    model = DecodeModel(tiny("awq"), quant_policy=8, max_batch=8, device="cpu")
    model.load_synthetic(seed=0)
    eng = TurboMind(tiny("awq"), TurbomindEngineConfig(max_batch_size=8, session_len=256, cache_max_entry_count=32, quant_policy=8),
                    model=model)
    assert eng.pool.num_blocks == 32 and eng.d_block_ptrs.numel() == 8 * 4
    inst = TurboMindInstance(eng)
    a = inst._submit(1, [5, 6, 7, 8, 9, 10], GenerationConfig(max_new_tokens=6, ignore_eos=True))
    b = inst._submit(2, [11, 12], GenerationConfig(max_new_tokens=7, ignore_eos=True))
    n0 = len(dry.calls)
    eng.step()                                   # a: 4 prompt tokens, b: 2 -> one mixed step of 6 rows
    name, args = [c for c in dry.calls[n0:] if c[0].startswith("b200_model_")][-1]
    assert name == "b200_model_chunk_step" and args[2] == 6 and args[1]._obj.batch_size == 2
    assert eng.d_tokens[:6].tolist() == [5, 6, 7, 8, 11, 12] and eng.d_cu_q[:3].tolist() == [0, 4, 6]
    assert eng.d_cu_k[:3].tolist() == [0, 4, 6] and eng.d_cu_blocks[:3].tolist() == [0, 1, 2]
    assert eng.d_block_ptrs[:2].tolist() == [eng.pool.ptr(a.blocks[0]), eng.pool.ptr(b.blocks[0])]
    assert b.generated == [0] and a.generated == []            # the recorder "samples" token 0 for every row
    eng.step()                                   # a: its last 2 prompt tokens (chunk), b: decode -> mixed again
    assert dry.calls[-1][0] == "b200_model_chunk_step" or "b200_model_chunk_step" in dry.names()[-3:]
    n1 = len(dry.calls)
    eng.step(), eng.step()                       # both decoding: eager the first time that batch size is seen, then a graph
    names = dry.names()[n1:]
    assert names.count("b200_model_decode_step") == 2 and names.count("graph.replay") == 1
    while eng.scheduler.has_work:
        eng.step()
    assert a.finish_reason == b.finish_reason == "length" and (len(a.generated), len(b.generated)) == (6, 7)
    m = eng.get_schedule_metrics()
    assert m.active_seqs == 0 and m.free_blocks == 32


@pytest.mark.parametrize("ingest_limit", [150000, 0])
def test_bench_main_arm_prints_the_contract_line(dry, monkeypatch, capsys, ingest_limit):
    """bench.run_b200 end to end on a tiny geometry with the recording library: every key the driver reads is on the one
    JSON line, with the right types -- the script itself is what the round-end run executes."""
    import json
    import os
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench
    import lmdeploy_b200.turbomind as TM
    from lmdeploy_b200.turbomind.runtime import DecodeModel

    class CpuModel(DecodeModel):
        def __init__(self, spec, **kw):
            super().__init__(spec, device="cpu", **kw)

    class Event:
        def __init__(self, enable_timing=False):
            pass

        def record(self, *a):
            pass

        def elapsed_time(self, other):
            return 2.0
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(TM, "DecodeModel", CpuModel)
    monkeypatch.setitem(TM.SPECS, "tiny", tiny("awq"))
    monkeypatch.setitem(bench.CONFIGS, "c2", ("tiny", 4, 100, 8))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    args = types.SimpleNamespace(gpus=1, steps=4, warmup=3, impl="b200", config="c2", no_graph=False, no_cpu_baseline=False,
                                 attn_iters=2, engine_ingest_limit=ingest_limit)
    bench.run_b200(args)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["value"] == pytest.approx(4 * 4 / 2e-3) and d["ms_per_step"] == pytest.approx(0.5)
    assert d["unit"] == "tokens/s" and d["dtype"] == "f16" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
    assert d["roofline"]["frac"] == pytest.approx(d["roofline"]["achieved"] / d["roofline"]["peak"])
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] >= (4 * 4 + 3) * 4     # the packed control buffer, + the block table when it changes
    assert "TurboMind.step" in d["e2e"]["path"] and d["e2e"]["prompt_ingest"]["tokens"] == 4 * 100
    assert d["e2e"]["prompt_ingest"]["how"].startswith("engine" if ingest_limit else "synthetic")
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["gpu_launches"] == (4 + 9 * 2 + 4) * 4 and "workload" in d["config"] and "model" not in d["config"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the timed loops ran through the captured graph: warm-up + timed + host-fed steps
    assert dry.names().count("graph.replay") >= 3 + 4 + 3 + 4


@pytest.mark.parametrize("arch,family", [("internlm2", "awq"), ("llama", None), ("llama", "gptq"), ("llama", "compressed-tensors"),
                                         ("llama", "fp8"), ("qwen2", "awq"), ("qwen3", None)])
@pytest.mark.parametrize("tp", [1, 2])
def test_checkpoint_directory_to_c_abi_upload(dry, monkeypatch, tmp_path, arch, family, tp):
    """load_hf_checkpoint end to end without a device: config.json + safetensors -> format resolver -> layer conversion
    -> Lin.to_tm -> DecodeModel.load_tm_weights -> b200_linear_prepare / b200_model_set_*.  Checks what the C ABI is
    handed: formats, dims of this TP rank, dtypes and shapes of the buffers (the recorder cannot see device memory, so the
    tensors are intercepted one level up)."""
    import json
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from safetensors.torch import save_file
    from test_loader_cpu import D, H, HKV, HQ, I, V, make_checkpoint
    import lmdeploy_b200.turbomind.runtime as RT
    from lmdeploy_b200.turbomind import loader as LD
    cfg, t, _ = make_checkpoint(arch, family, seed=3)
    (tmp_path / "config.json").write_text(json.dumps(dict(cfg, eos_token_id=[2, 7])))
    save_file({k: v.contiguous() for k, v in t.items()}, str(tmp_path / "model.safetensors"))
    seen = []
    real_set_linear = RT.DecodeModel._set_linear

    class CpuModel(RT.DecodeModel):
        def __init__(self, spec, **kw):
            super().__init__(spec, device="cpu", **kw)

        def _set_linear(self, layer, which, K, N, weight, scales=None, zeros=None, transposed=False, fmt=None):
            seen.append((layer, which, K, N, weight, scales, zeros, fmt))
            return real_set_linear(self, layer, which, K, N, weight, scales, zeros, transposed, fmt)
    monkeypatch.setattr(RT, "DecodeModel", CpuModel)
    spec, model, eos = LD.load_hf_checkpoint(str(tmp_path), quant_policy=8, max_batch=8, tp=tp, rank=tp - 1)
    assert eos == 2 and spec.arch == arch and model.tp_size == tp
    hq, hkv, it = HQ // tp, LD.kv_heads_for_tp(HKV, tp) // tp, I // tp
    want_dims = {L.W_QKV: (H, (hq + 2 * hkv) * D), L.W_O: (hq * D, H), L.W_GATE_UP: (H, 2 * it), L.W_DOWN: (it, H)}
    int4 = family in ("awq", "gptq", "compressed-tensors")
    T = torch.bfloat16 if family == "fp8" else torch.float16
    per_layer = [s for s in seen if s[1] != L.W_LM_HEAD]
    assert len(per_layer) == 4 * spec.layers and len(seen) == len(per_layer) + 1
    for layer, which, K, N, w, sc, z, fmt in per_layer:
        assert (K, N) == want_dims[which], (which, K, N)
        if int4:
            assert w.dtype == torch.int32 and tuple(w.shape) == (K, N // 8) and fmt is None
            assert sc.dtype == z.dtype == T and tuple(sc.shape) == tuple(z.shape) == (K // 128, N)
        elif family == "fp8":
            assert w.dtype == torch.uint8 and tuple(w.shape) == (K, N) and fmt == L.FMT_FP8_COL
            assert sc.dtype == torch.float32 and tuple(sc.shape) == (K // 128, N) and z is None
        else:
            assert w.dtype == T and tuple(w.shape) == (K, N) and sc is None and z is None
        assert w.is_contiguous()
    lm = seen[-1]
    assert lm[1] == L.W_LM_HEAD and (lm[2], lm[3]) == (H, (V + 128 * tp - 1) // (128 * tp) * 128) and lm[4].dtype == T
    fmts = [a[0]._obj.weight_format for n, a in dry.calls if n == "b200_linear_prepare"]
    assert fmts[:-1] == [L.FMT_U4 if int4 else L.FMT_FP8_COL if family == "fp8" else L.FMT_DENSE] * (4 * spec.layers)
    ids = [a[2] for n, a in dry.calls if n == "b200_model_set_tensor"]
    assert ids.count(L.T_QKV_BIAS) == (spec.layers if arch == "qwen2" else 0)
    assert ids.count(L.T_Q_NORM) == ids.count(L.T_K_NORM) == (spec.layers if arch == "qwen3" else 0)
    cfgc = [a[1]._obj for n, a in dry.calls if n == "b200_model_create"][-1]
    assert (cfgc.num_heads, cfgc.num_kv_heads, cfgc.inter_size, cfgc.tp_size, cfgc.tp_rank) == (hq, hkv, it, tp, tp - 1)
    assert cfgc.dtype == (1 if family == "fp8" else 0) and cfgc.vocab == lm[3] and cfgc.embedding_rows == V


def test_pipeline_over_a_checkpoint_directory(dry, monkeypatch, tmp_path):
    """lmdeploy_b200.pipeline(<dir>) -> TurboMind.from_pretrained -> load_hf_checkpoint (hf_overrides applied) -> engine ->
    infer, without a device: the recorder's step leaves token 0 in the output buffer, so every sampled id is 0."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from safetensors.torch import save_file
    from test_loader_cpu import make_checkpoint
    import lmdeploy_b200
    import lmdeploy_b200.turbomind.runtime as RT
    cfg, t, _ = make_checkpoint("llama", "awq", seed=1)
    (tmp_path / "config.json").write_text(json.dumps(dict(cfg, max_position_embeddings=96)))
    save_file({k: v.contiguous() for k, v in t.items()}, str(tmp_path / "model.safetensors"))

    class CpuModel(RT.DecodeModel):
        def __init__(self, spec, **kw):
            super().__init__(spec, device="cpu", **kw)
    monkeypatch.setattr(RT, "DecodeModel", CpuModel)
    ecfg = lmdeploy_b200.TurbomindEngineConfig(max_batch_size=8, cache_max_entry_count=16, quant_policy=8,
                                               hf_overrides={"rope_scaling": {"type": "dynamic", "factor": 2.0}})
    with lmdeploy_b200.pipeline(str(tmp_path), ecfg) as pipe:
        eng = pipe.engine
        assert eng.spec.rope.type == "dynamic" and eng.session_len == 96 and eng.max_batch == 8 and eng.pool.num_blocks == 16
        r = pipe([[5, 6, 7]], lmdeploy_b200.GenerationConfig(max_new_tokens=3))[0]
        assert (r.finish_reason, r.token_ids, r.input_token_len, r.generate_token_len) == ("length", [0, 0, 0], 3, 3)
        # eos ids of generation_config.json end the turn and stay out of the response
        pipe.hf_gen_cfg = {"eos_token_id": [0, 9]}
        r = pipe([5, 6, 7], lmdeploy_b200.GenerationConfig(max_new_tokens=3))
        assert (r.finish_reason, r.token_ids, r.generate_token_len) == ("stop", [], 0)
        # a prompt beyond the trained context would need a per-request rope base under dynamic NTK: refused
        with pytest.raises(NotImplementedError):
            pipe([list(range(97))], lmdeploy_b200.GenerationConfig(max_new_tokens=1))
    assert pipe.engine is None


def test_model_format_must_agree_with_the_checkpoint(dry, monkeypatch, tmp_path):
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from safetensors.torch import save_file
    from test_loader_cpu import make_checkpoint
    import lmdeploy_b200.turbomind.runtime as RT
    from lmdeploy_b200.turbomind import loader as LD
    cfg, t, _ = make_checkpoint("llama", "awq", seed=1)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    save_file({k: v.contiguous() for k, v in t.items()}, str(tmp_path / "model.safetensors"))

    class CpuModel(RT.DecodeModel):
        def __init__(self, spec, **kw):
            super().__init__(spec, device="cpu", **kw)
    monkeypatch.setattr(RT, "DecodeModel", CpuModel)
    with pytest.raises(ValueError, match="mismatched quant method"):
        LD.load_hf_checkpoint(str(tmp_path), model_format="gptq")
    spec, _, _ = LD.load_hf_checkpoint(str(tmp_path), model_format="awq", max_batch=4)
    assert spec.checkpoint_format == "awq"


def test_logits_and_logprobs_outputs(dry, monkeypatch):
    """output_logits='generation' / logprobs=k: the engine reads the row each accepted token was sampled from out of the
    step's logits buffer (stubbed here) -- fp32 logits over the REAL vocabulary, {id: logprob} of the top k plus the
    sampled id; the pipeline drops the entries of a hidden stop token together with the token."""
    from lmdeploy_b200.api import Pipeline
    from lmdeploy_b200.messages import GenerationConfig, TurbomindEngineConfig
    from lmdeploy_b200.turbomind.engine import TurboMind, TurboMindInstance
    from lmdeploy_b200.turbomind.runtime import DecodeModel
    monkeypatch.setenv("B200_CHUNKED_PREFILL", "1")     # one row per sequence per step: sequence i samples from row i
    model = DecodeModel(tiny("awq"), quant_policy=8, max_batch=8, device="cpu")
    model.load_synthetic(seed=0)
    eng = TurboMind(tiny("awq"), TurbomindEngineConfig(max_batch_size=8, session_len=64, cache_max_entry_count=16, quant_policy=8),
                    model=model)
    buf = (torch.arange(model.vocab_local, dtype=torch.float32)[None] * 0.01 + torch.arange(8, dtype=torch.float32)[:, None]).half()
    monkeypatch.setattr(model, "logits", lambda n: buf[:n])
    p = Pipeline.__new__(Pipeline)
    p.engine, p.tokenizer = eng, None
    cfg = GenerationConfig(max_new_tokens=3, ignore_eos=True, output_logits="generation", logprobs=2)
    r, plain = p.infer([[5, 6, 7], [8, 9]], [cfg, GenerationConfig(max_new_tokens=2, ignore_eos=True)])
    assert plain.logits is None and plain.logprobs is None
    assert tuple(r.logits.shape) == (3, 1000) and r.logits.dtype == torch.float32          # the real vocabulary, not the padded one
    assert torch.equal(r.logits[0], buf[0, :1000].float())                                 # sequence 0 samples from row 0
    assert len(r.logprobs) == 3 and all(set(d) == {999, 998, 0} for d in r.logprobs)       # top-2 of an increasing row + sampled id 0
    want = torch.log_softmax(buf[0, :1000].float(), -1)
    assert r.logprobs[0][999] == pytest.approx(float(want[999])) and r.logprobs[0][0] == pytest.approx(float(want[0]))
    # streamed: every content chunk carries the entries of its own tokens
    chunks = list(p.stream_infer([[5, 6, 7]], gen_config=cfg))
    assert [len(c.logprobs) for c in chunks[:-1]] == [1, 1, 1] and all(tuple(c.logits.shape) == (1, 1000) for c in chunks[:-1])
    # a stop token is hidden with its entries (the recorder samples id 0, declared a stop id here)
    r = p.infer([[5, 6, 7]], GenerationConfig(max_new_tokens=3, stop_token_ids=[0], output_logits="generation", logprobs=1))[0]
    assert (r.finish_reason, r.token_ids, r.logprobs, r.logits) == ("stop", [], [], None)
    # engine-level outputs, and the refusals
    outs = list(TurboMindInstance(eng).stream_infer(77, [5, 6], cfg))
    assert [len(o.logprobs) for o in outs] == [1, 1, 1] or sum(len(o.logprobs) for o in outs) == 3
    with pytest.raises(NotImplementedError):
        GenerationConfig(output_logits="all")
    assert GenerationConfig(logprobs=5000).logprobs == 1024
    eng.channel = object()
    with pytest.raises(NotImplementedError):
        TurboMindInstance(eng)._submit(78, [1, 2], cfg)


def test_do_sample_requests_go_through_the_device_sampling_stage(dry, monkeypatch):
    """A do_sample request: the engine seeds one RNG slot for it, calls b200_sample for its row with its parameters every step it
    produces a token (the recording library "samples" id 0), greedy requests of the same batch keep the arg-max; the slot
    returns to the pool when the request ends."""
    import ctypes as C
    from lmdeploy_b200.messages import GenerationConfig, TurbomindEngineConfig
    from lmdeploy_b200.turbomind.engine import TurboMind, TurboMindInstance
    from lmdeploy_b200.turbomind.runtime import DecodeModel
    monkeypatch.setenv("B200_CHUNKED_PREFILL", "1")
    model = DecodeModel(tiny("awq"), quant_policy=8, max_batch=8, device="cpu")
    model.load_synthetic(seed=0)
    eng = TurboMind(tiny("awq"), TurbomindEngineConfig(max_batch_size=8, session_len=64, cache_max_entry_count=16, quant_policy=8),
                    model=model)
    inst = TurboMindInstance(eng)
    a = inst._submit(1, [5, 6], GenerationConfig(max_new_tokens=3, ignore_eos=True, do_sample=True, top_k=40, top_p=0.9,
                                                 temperature=0.7, random_seed=123, repetition_penalty=1.2))
    b = inst._submit(2, [7], GenerationConfig(max_new_tokens=3, ignore_eos=True))
    assert a.sampling == dict(temperature=0.7, top_k=40, top_p=0.9, min_p=0.0, repetition_penalty=1.2, seed=123) and b.sampling is None
    while eng.scheduler.has_work:
        eng.step()
    names = dry.names()
    assert names.count("b200_sampling_init_rng") == 1 and names.count("b200_sample") == 3      # one per generated token of `a`
    p = [args for n, args in dry.calls if n == "b200_sample"][0][0]._obj
    assert p.count == 1 and p.vocab == 1000 and p.ld == model.vocab_local and p.repetition_penalty and p.history
    assert a.rng_slot >= 0 and len(a.generated) == 3 and len(b.generated) == 3


def test_tp_group_picks_the_exchange_by_world_size(dry, monkeypatch):
    """TpGroup: CUDA-IPC peer loads below 4 ranks, NVLS (torch symmetric memory + multicast) from 4 ranks up with a silent
    fall-back to peer loads when the fabric offers none, B200_TP_NVLS=1 insists, =0 forbids (measured: profiles/r02_scale.md)."""
    import types
    import warnings
    import torch.distributed as dist
    from lmdeploy_b200.turbomind import tp as TP
    monkeypatch.setattr(dist, "all_gather_object", lambda out, obj, group=None: out.__setitem__(slice(None), [obj] * len(out)))
    monkeypatch.setattr(dist, "barrier", lambda group=None: None)
    tried = []

    def no_multicast(self, model, rank, world, nbytes, fbytes, group):
        tried.append(world)
        raise RuntimeError("no multicast on this fabric")
    monkeypatch.setattr(TP.TpGroup, "_init_symmetric", no_multicast)
    model = types.SimpleNamespace(spec=types.SimpleNamespace(hidden=256), max_batch=8, vocab_local=128,
                                  device=torch.device("cpu"), set_comm=lambda h: None)
    monkeypatch.delenv("B200_TP_NVLS", raising=False)
    g = TP.TpGroup(model, 0, 2)
    assert tried == [] and not g.nvls and "b200_comm_create" in dry.names()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        g = TP.TpGroup(model, 1, 4)
    assert tried == [4] and not g.nvls and any("NVLS exchange unavailable" in str(x.message) for x in w)
    monkeypatch.setenv("B200_TP_NVLS", "0")
    TP.TpGroup(model, 0, 8)
    assert tried == [4]
    monkeypatch.setenv("B200_TP_NVLS", "1")
    import pytest
    with pytest.raises(RuntimeError, match="no multicast"):
        TP.TpGroup(model, 0, 2)
