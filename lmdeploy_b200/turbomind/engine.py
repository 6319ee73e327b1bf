"""TurboMind / TurboMindInstance: the host-side engine the Python API drives.

Mirrors lmdeploy/turbomind/turbomind.py (TurboMind :108-566, TurboMindInstance :568-813): an engine object owning
weights + KV pool, instances that stream tokens for a session.  The model forward is the native decode step
(libb200tm.so); scheduling is lmdeploy_b200.turbomind.scheduler."""
import asyncio
import os
import threading
from typing import List, Optional

import torch

from ..messages import DEFAULT_MAX_BATCH_SIZE, EngineOutput, GenerationConfig, ResponseType, TurbomindEngineConfig
from .runtime import DecodeModel, KvCachePool
from .scheduler import InputLengthError, Scheduler, Sequence
from .specs import SPECS, ModelSpec


def marshal_step(batch, block_ptr=None):
    """Argument arrays of one (mixed) step for `batch` = [Sequence]: sequence i feeds its `step_tokens` next tokens as
    rows [cu_q_len[i], cu_q_len[i+1]); cu_k_len counts cached history + the fed tokens (AttentionParams::cu_k_len
    includes the new tokens); block_ids is the concatenation of the sequences' block tables (pool indices -- what a TP
    leader broadcasts; every rank maps them to addresses of its own pool) and block_ptrs their addresses through
    `block_ptr` when one is given."""
    tokens, cu_q, cu_k, ids, cu_blocks = [], [0], [0], [], [0]
    q = k = max_k = 0
    for s in batch:
        n = s.step_tokens
        cached = s.cached
        prompt = s.prompt
        if n == 1 and cached < len(prompt) + len(s.generated):      # steady decode: one row, the token at position `cached`
            tokens.append(int(prompt[cached] if cached < len(prompt) else s.generated[cached - len(prompt)]))
        else:
            toks = s.next_inputs()
            if len(toks) != n:
                raise RuntimeError(f"session {s.session_id}: scheduled {n} tokens but {len(toks)} are available")
            tokens += [int(t) for t in toks]
        blocks = s.blocks
        if not blocks:
            raise RuntimeError(f"session {s.session_id} runs without KV blocks")
        q += n
        end = cached + n
        k += end
        if end > max_k:
            max_k = end
        cu_q.append(q)
        cu_k.append(k)
        ids += blocks
        cu_blocks.append(len(ids))
    st = dict(batch_size=len(batch), token_ids=tokens, token_num=len(tokens), cu_q_len=cu_q, cu_k_len=cu_k, block_ids=ids,
              cu_block_nums=cu_blocks, max_k_len=max_k)
    if block_ptr is not None:
        st["block_ptrs"] = [block_ptr(b) for b in ids]
    return st


class TurboMind:
    def __init__(self, spec: ModelSpec, engine_config: Optional[TurbomindEngineConfig] = None, *, model: DecodeModel = None,
                 eos_id: Optional[int] = None, num_blocks: Optional[int] = None, seed=0, channel=None):
        cfg = engine_config or TurbomindEngineConfig()
        if cfg.tp != 1 and (model is None or channel is None):
            raise NotImplementedError("tp > 1: launch one process per GPU (torchrun), give every rank its TP-sharded "
                                      "DecodeModel (lmdeploy_b200.turbomind.tp.TpGroup) and a StepChannel "
                                      "(lmdeploy_b200.turbomind.tp_engine); rank 0 drives step(), the others "
                                      "serve_follower()")
        self.channel = channel
        self.config = cfg
        self.spec = spec
        self.eos_id = eos_id
        self.max_batch = cfg.max_batch_size or DEFAULT_MAX_BATCH_SIZE
        # session_len left unset: what _get_and_verify_max_len derives from the HF config (converter.py get_tm_config step 4;
        # loader.default_session_len), 2048 when the config names no context length
        self.session_len = cfg.session_len or spec.session_len or spec.max_position_embeddings or 2048
        self.block_len = cfg.cache_block_seq_len
        self.model = model or DecodeModel(spec, quant_policy=cfg.quant_policy, block_len=self.block_len,
                                          max_batch=self.max_batch)
        if model is None:
            self.model.load_synthetic(seed)
        dev = self.model.device
        if num_blocks is None:
            if cfg.cache_max_entry_count >= 1:
                num_blocks = int(cfg.cache_max_entry_count)
            else:   # ratio of free memory, as CreateEngine does (turbomind.cc:293-311)
                free, _ = torch.cuda.mem_get_info(dev)
                num_blocks = int(free * cfg.cache_max_entry_count) // (self.model.layer_bytes * spec.layers)
        self.pool = KvCachePool(num_blocks, spec.layers, self.model.kv_heads, spec.head_dim, self.block_len,
                                cfg.quant_policy, device=dev)
        # Chunked prefill: a sequence still inside its prompt feeds up to B200_CHUNKED_PREFILL tokens (default 64) per step
        # through b200_model_chunk_step (tcgen05 prefill attention; parity-green on a B200 against the oracle, the
        # reference's own prefill kernels and token-by-token feeding: tests/test_prefill_attention_gpu.py,
        # tests/test_engine_gpu.py) as long as the step's row budget (max_batch rows) has room.  1 = one prompt token per
        # step through the decode kernel.
        chunk = int(os.environ.get("B200_CHUNKED_PREFILL", "64"))
        self.scheduler = Scheduler(max_batch=self.max_batch, block_len=self.block_len, num_blocks=num_blocks,
                                   session_len=self.session_len, prefix_caching=cfg.enable_prefix_caching,
                                   prefill_chunk=min(chunk, cfg.max_prefill_token_num), token_budget=self.max_batch)
        B = self.max_batch
        self._max_blocks = (self.session_len + self.block_len - 1) // self.block_len
        # Per-step control arrays live in ONE int32 device buffer [tokens B | cu_q B+1 | cu_k B+1 | cu_blocks B+1] fed by one
        # pinned staging buffer: a step costs a single ~1 KB H2D copy (plus the block table only when it changed) --
        # the counterpart of the pinned staging in UnifiedAttentionLayer::Setup (unified_attention_layer.cc:288-305).
        self._o_q, self._o_k, self._o_b, n_ctl = B, 2 * B + 1, 3 * B + 2, 4 * B + 3
        self.d_ctl = torch.zeros(n_ctl, dtype=torch.int32, device=dev)
        self.h_ctl = torch.zeros(n_ctl, dtype=torch.int32).pin_memory()
        self._h_ctl = self.h_ctl.numpy()
        self.d_tokens = self.d_ctl[:B]
        self.d_cu_q = self.d_ctl[self._o_q:self._o_q + B + 1]
        self.d_cu_k = self.d_ctl[self._o_k:self._o_k + B + 1]
        self.d_cu_blocks = self.d_ctl[self._o_b:self._o_b + B + 1]
        self.d_out = torch.zeros(B, dtype=torch.int32, device=dev)
        self.d_block_ptrs = torch.zeros(B * self._max_blocks, dtype=torch.int64, device=dev)
        self.h_out = torch.zeros(B, dtype=torch.int32).pin_memory()
        self._h_out = self.h_out.numpy()
        self.h_block_ptrs = torch.zeros(B * self._max_blocks, dtype=torch.int64).pin_memory()
        self._h_block_ptrs = self.h_block_ptrs.numpy()
        self._last_block_ids = None
        self.h2d_bytes = self.d2h_bytes = 0          # running totals of the step traffic (bench.py reports them per step)
        self._lock = threading.Lock()
        self._next_session = 0
        self._use_graphs = os.environ.get("B200_ENGINE_GRAPHS", "1") != "0"
        self._graphs, self._seen_sizes = {}, set()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, engine_config=None, **kw):
        """`pretrained_model_name_or_path`: a key of SPECS (synthetic random weights of that geometry; no real
        checkpoints exist in this image) or a directory holding an HF-format checkpoint (config.json +
        *.safetensors, InternLM2 / Llama, fp16 or AWQ) read by lmdeploy_b200.turbomind.loader."""
        import os
        if pretrained_model_name_or_path in SPECS:
            return cls(SPECS[pretrained_model_name_or_path], engine_config, **kw)
        if os.path.isdir(pretrained_model_name_or_path):
            from .loader import load_hf_checkpoint
            cfg = engine_config or TurbomindEngineConfig()
            spec, model, eos = load_hf_checkpoint(pretrained_model_name_or_path, quant_policy=cfg.quant_policy,
                                                  block_len=cfg.cache_block_seq_len, hf_overrides=cfg.hf_overrides,
                                                  model_format=cfg.model_format,
                                                  max_batch=cfg.max_batch_size or DEFAULT_MAX_BATCH_SIZE)
            return cls(spec, cfg, model=model, eos_id=eos, **kw)
        raise FileNotFoundError(pretrained_model_name_or_path)

    def create_instance(self):
        return TurboMindInstance(self)

    # ------------------------------------------------------------------ one engine step
    def step(self):
        """Schedule + run one step for the active batch; returns [(sequence, token|None)] events.  Under tensor
        parallelism only the leader (rank 0) calls this; it publishes the step's arrays to the followers first."""
        sched = self.scheduler
        batch = sched.schedule()
        if not batch:
            # a tick may legitimately form no batch (it only retired or preempted); a long run of them with work pending
            # means the scheduler cannot make progress -- fail loudly instead of spinning in the callers' loops
            self._idle_ticks = getattr(self, "_idle_ticks", 0) + 1
            if sched.has_work and self._idle_ticks > 1000:
                raise RuntimeError(f"scheduler stalled: {sched.metrics()}")
            return []
        self._idle_ticks = 0
        st = marshal_step(batch)
        if self.channel is not None:
            self.channel.publish(st)
        out = self._execute(st)
        cu_q = st["cu_q_len"]
        rows = [cu_q[i + 1] - 1 for i in range(len(batch))]                              # sample = last row of each sequence
        sampled = [out[r] for r in rows]
        if any(s.sampling is not None for s in batch):
            self._sample(batch, rows, sampled)
        events = sched.update(batch, sampled)
        if any(s.want_logits or s.want_logprobs for s in batch):
            self._collect_extras(batch, rows, events, st["token_num"])
        return events

    def _sample(self, batch, rows, sampled):
        """do_sample sequences that produce a token in this step: the device sampling stage (b200_sample: repetition penalty,
        temperature, top-k, top-p, min-p, one XORWOW state per request seeded by GenerationConfig.random_seed -- what
        LogitsProcessor + Sampling do in the reference, generation/logits_processor.cc:66-125, generation/sampling.cc:92-183)
        replaces the step's arg-max for their rows, reading the same logits buffer."""
        import ctypes as C
        import numpy as np
        from .. import _lib as L
        lib, dev = self.model.lib, self.model.device
        items = [(i, s) for i, s in enumerate(batch) if s.sampling is not None and s.cached + s.step_tokens >= s.context_len]
        if not items:
            return
        if not hasattr(self, "_rng_states"):
            self._rng_bytes = lib.b200_sampling_rng_state_bytes()
            self._rng_states = torch.zeros(self.max_batch * self._rng_bytes, dtype=torch.uint8, device=dev)
            self._rng_free = list(range(self.max_batch - 1, -1, -1))
            self._samp_ws = torch.zeros(lib.b200_sampling_workspace_bytes(self.max_batch, self.spec.vocab), dtype=torch.uint8,
                                        device=dev)
            self._rng_owner = {}
        stream = torch.cuda.current_stream().cuda_stream
        # slots of finished requests return to the free list; new requests get a slot and their seed
        for slot, owner in list(self._rng_owner.items()):
            if owner.finished:
                del self._rng_owner[slot]
                self._rng_free.append(slot)
        fresh = [s for _, s in items if s.rng_slot < 0]
        if fresh:
            seeds = np.zeros(self.max_batch, np.uint64)
            mask = np.zeros(self.max_batch, np.uint8)
            for s in fresh:
                s.rng_slot = self._rng_free.pop()
                self._rng_owner[s.rng_slot] = s
                seeds[s.rng_slot], mask[s.rng_slot] = s.sampling["seed"] & 0xFFFFFFFFFFFFFFFF, 1
            d_seeds = torch.from_numpy(seeds.view(np.int64)).to(dev)
            d_mask = torch.from_numpy(mask).to(dev)
            L.check(lib.b200_sampling_init_rng(self._rng_states.data_ptr(), d_seeds.data_ptr(), d_mask.data_ptr(), self.max_batch,
                                               stream), "b200_sampling_init_rng")
        n = len(items)
        f32 = lambda key: torch.tensor([s.sampling[key] for _, s in items], dtype=torch.float32, device=dev)
        d_rows = torch.tensor([rows[i] for i, _ in items], dtype=torch.int32, device=dev)
        d_t, d_p, d_m, d_pen = f32("temperature"), f32("top_p"), f32("min_p"), f32("repetition_penalty")
        d_k = torch.tensor([s.sampling["top_k"] for _, s in items], dtype=torch.int32, device=dev)
        d_slot = torch.tensor([s.rng_slot for _, s in items], dtype=torch.int32, device=dev)
        d_out = torch.zeros(n, dtype=torch.int32, device=dev)
        p = L.SamplingParams()
        p.logits, p.ld = lib.b200_model_buffer(self.model.handle, 0), self.model.vocab_local
        p.dtype, p.count, p.vocab = self.model.cfg.dtype, n, self.spec.vocab
        p.rows, p.temperature, p.top_k, p.top_p, p.min_p = (d_rows.data_ptr(), d_t.data_ptr(), d_k.data_ptr(), d_p.data_ptr(),
                                                            d_m.data_ptr())
        hist = None
        if any(s.sampling["repetition_penalty"] != 1.0 for _, s in items):
            # the ids seen so far (prompt + generated), uploaded for the requests that ask for the penalty
            hist = [torch.tensor((s.prompt + s.generated) if s.sampling["repetition_penalty"] != 1.0 else [0], dtype=torch.int32,
                                 device=dev) for _, s in items]
            d_hp = torch.tensor([h.data_ptr() for h in hist], dtype=torch.int64, device=dev)
            d_hl = torch.tensor([h.numel() if s.sampling["repetition_penalty"] != 1.0 else 0 for h, (_, s) in zip(hist, items)],
                                dtype=torch.int32, device=dev)
            p.repetition_penalty, p.history, p.history_len = d_pen.data_ptr(), d_hp.data_ptr(), d_hl.data_ptr()
        p.rng_states, p.rng_index, p.out_ids = self._rng_states.data_ptr(), d_slot.data_ptr(), d_out.data_ptr()
        p.workspace, p.workspace_bytes, p.stream = self._samp_ws.data_ptr(), self._samp_ws.numel(), stream
        L.check(lib.b200_sample(C.byref(p)), "b200_sample")
        ids = d_out.cpu().tolist()
        for (i, _), t in zip(items, ids):
            sampled[i] = int(t)

    def _collect_extras(self, batch, rows, events, token_num):
        """output_logits='generation' / logprobs=k: for every token accepted in this step, the fp32 logits of the row it was
        sampled from and/or {token id: log-probability} of the k most likely ids plus the sampled one -- what the
        reference returns (turbomind.py _get_logits / _get_logprobs).  Read from the step's logits buffer (T, one GPU)."""
        row_of = {id(s): r for s, r in zip(batch, rows)}
        lg = None
        for s, tok in events:
            if tok is None or not (s.want_logits or s.want_logprobs):
                continue
            if lg is None:
                lg = self.model.logits(token_num)[:, :self.spec.vocab].float()
            row = lg[row_of[id(s)]]
            if s.want_logits:
                s.logits.append(row.cpu())
            if s.want_logprobs:
                lp = torch.log_softmax(row, -1)
                top = torch.topk(lp, min(s.want_logprobs, lp.numel()))
                d = {int(i): float(v) for v, i in zip(top.values.tolist(), top.indices.tolist())}
                d[int(tok)] = float(lp[int(tok)])
                s.logprobs.append(d)

    def serve_follower(self):
        """Tensor-parallel rank > 0: execute the steps the leader publishes until it shuts the channel down.  Every rank
        owns a KV pool of the same geometry, so the leader's block ids are valid here; the sampled ids are the same on
        every rank (vocab-sharded argmax exchange) and are ignored by the followers."""
        while True:
            st = self.channel.receive()
            if st is None:
                return
            self._execute(st)

    def _execute(self, st):
        """Upload one step's arrays, run it on this rank's model shard, return the sampled id of every row."""
        B, T = st["batch_size"], st["token_num"]
        h = self._h_ctl
        h[:T] = st["token_ids"]
        h[self._o_q:self._o_q + B + 1] = st["cu_q_len"]
        h[self._o_k:self._o_k + B + 1] = st["cu_k_len"]
        h[self._o_b:self._o_b + B + 1] = st["cu_block_nums"]
        self.d_ctl.copy_(self.h_ctl, non_blocking=True)
        self.h2d_bytes += self.h_ctl.numel() * 4
        ids = st["block_ids"]
        if ids != self._last_block_ids:          # the block table changes only when a sequence crosses a block boundary
            n = len(ids)
            hp = self._h_block_ptrs[:n]
            hp[:] = ids
            hp *= self.pool.block_bytes
            hp += self.pool.base
            self.d_block_ptrs[:n].copy_(self.h_block_ptrs[:n], non_blocking=True)
            self.h2d_bytes += n * 8
            self._last_block_ids = ids

        def launch():
            self.model.decode_step(batch=B, token_ids=self.d_tokens, cu_q_len=self.d_cu_q, cu_k_len=self.d_cu_k,
                                   block_ptrs=self.d_block_ptrs, cu_block_nums=self.d_cu_blocks, out_tokens=self.d_out,
                                   max_k_len=st["max_k_len"], token_num=T)

        # A pure decode step is shape-static given the batch size (every argument lives in a fixed device buffer, the
        # attention work partition is computed on the device), so its ~300 launches are replayed from a CUDA graph per
        # batch size -- captured the second time that size is seen (the first run doubles as the warm-up).
        if T == B and self._use_graphs:
            g = self._graphs.get(B)
            if g is None and B in self._seen_sizes:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    launch()
                self._graphs[B] = g
            if g is not None:
                g.replay()
            else:
                launch()
                self._seen_sizes.add(B)
        else:
            launch()
        self.h_out.copy_(self.d_out, non_blocking=True)
        self.d2h_bytes += self.h_out.numel() * 4
        torch.cuda.current_stream().synchronize()
        return self._h_out[:T].tolist()

    def close(self):
        """Release tensor-parallel followers (if any) and the communicator."""
        if self.channel is not None and self.channel.is_leader:
            self.channel.shutdown()
        tp = getattr(self, "_tp", None)
        if tp is not None:
            tp.close()
            self._tp = None

    def get_schedule_metrics(self, index=0):
        """ScheduleMetrics of this rank (TurboMind.get_schedule_metrics, lmdeploy/turbomind/turbomind.py)."""
        from ..messages import ScheduleMetrics
        return ScheduleMetrics(**self.scheduler.metrics())

    def new_session_id(self):
        with self._lock:
            self._next_session += 1
            return self._next_session


class TurboMindInstance:
    """Stream tokens for one session.  Same call shape as the reference's
    TurboMindInstance.async_stream_infer(session_id, input_ids, gen_config=...) (turbomind.py:687-813)."""

    def __init__(self, engine: TurboMind):
        self.engine = engine

    def _submit(self, session_id, input_ids, gen_config):
        g = gen_config or GenerationConfig()
        # max_new_tokens = None: until the session is full (the scheduler clamps to session_len), as the reference does
        rope = getattr(getattr(self.engine, "spec", None), "rope", None)
        if rope is not None and rope.dynamic_ntk_base(len(input_ids)) != rope.base:
            raise NotImplementedError(f"a {len(input_ids)}-token prompt exceeds the model's trained context "
                                      f"({rope.max_position_embeddings}); dynamic-NTK rescaling of the rope base per "
                                      "request is not implemented on this path")
        budget = g.max_new_tokens if g.max_new_tokens is not None else 1 << 30
        seq = Sequence(session_id=session_id, prompt=list(map(int, input_ids)), max_new_tokens=budget,
                       stop_token_ids=list(g.stop_token_ids or []), ignore_eos=g.ignore_eos, eos_id=self.engine.eos_id)
        seq.want_logits, seq.want_logprobs = g.output_logits == "generation", int(g.logprobs or 0)
        if g.do_sample and not (g.top_k == 1 and g.repetition_penalty == 1.0):     # top_k == 1 without a penalty IS greedy
            if getattr(self.engine, "channel", None) is not None or self.engine.model.tp_size != 1:
                raise NotImplementedError("do_sample under tensor parallelism (the logits are vocabulary-sharded)")
            seed = g.random_seed if g.random_seed is not None else int.from_bytes(os.urandom(8), "little")
            seq.sampling = dict(temperature=float(g.temperature), top_k=int(g.top_k), top_p=float(g.top_p),
                                min_p=float(g.min_p), repetition_penalty=float(g.repetition_penalty), seed=int(seed))
        if (seq.want_logits or seq.want_logprobs) and getattr(self.engine, "channel", None) is not None:
            raise NotImplementedError("logits / logprobs outputs under tensor parallelism (the logits are vocabulary-sharded)")
        self.engine.scheduler.add(seq)
        return seq

    def stream_infer(self, session_id, input_ids, gen_config: GenerationConfig = None, **kw):
        try:
            seq = self._submit(session_id, input_ids, gen_config)
        except InputLengthError:
            yield EngineOutput(ResponseType.INPUT_LENGTH_ERROR, [], 0)
            return
        except KeyError:
            yield EngineOutput(ResponseType.SESSION_REPEAT, [], 0)
            return
        # EngineOutput.token_ids = "the newly generated token ids in each iteration" (lmdeploy/messages.py:721-743,
        # turbomind.py async_stream_infer: output_ids_buf[prev_len:seq_len]); a stop token is part of them.  num_token is
        # the running total.
        sent = 0
        while not seq.finished:
            self.engine.step()
            if len(seq.generated) > sent and not seq.finished:
                new, lo, sent = seq.generated[sent:], sent, len(seq.generated)
                yield EngineOutput(ResponseType.SUCCESS, list(new), sent, **self._extras(seq, lo, sent))
        yield EngineOutput(ResponseType.FINISH if seq.finish_reason != "cancel" else ResponseType.CANCEL,
                           list(seq.generated[sent:]), len(seq.generated), **self._extras(seq, sent, len(seq.generated)))

    @staticmethod
    def _extras(seq, lo, hi):
        out = {}
        if seq.want_logprobs:
            out["logprobs"] = seq.logprobs[lo:hi]
        if seq.want_logits and hi > lo:
            out["logits"] = torch.stack(seq.logits[lo:hi])
        return out

    async def async_stream_infer(self, session_id, input_ids, gen_config: GenerationConfig = None, **kw):
        for out in self.stream_infer(session_id, input_ids, gen_config):
            yield out
            await asyncio.sleep(0)

    def cancel(self, session_id):
        self.engine.scheduler.cancel(session_id)


def _per_prompt(gen_config, n):
    if isinstance(gen_config, (list, tuple)):
        if len(gen_config) != n:
            raise ValueError(f"{len(gen_config)} generation configs for {n} prompts")
        return list(gen_config)
    return [gen_config] * n


def _submit_or_reject(inst, engine, prompt, gen_config):
    """A prompt that cannot be admitted becomes a finished Sequence with finish_reason 'error' (the others still run)."""
    try:
        return inst._submit(engine.new_session_id(), prompt, gen_config)
    except InputLengthError as e:
        seq = Sequence(session_id=-1, prompt=list(map(int, prompt)), max_new_tokens=0, stop_token_ids=[], ignore_eos=True,
                       eos_id=None)
        seq.finished, seq.finish_reason, seq.error = True, "error", str(e)
        return seq


def batched_generate(engine, prompts: List[List[int]], gen_config=None):
    """Run a list of token-id prompts to completion with continuous batching; returns the Sequence of every prompt.
    gen_config: one GenerationConfig for all prompts or one per prompt."""
    inst = TurboMindInstance(engine)
    cfgs = _per_prompt(gen_config, len(prompts))
    seqs = [_submit_or_reject(inst, engine, p, g) for p, g in zip(prompts, cfgs)]
    while engine.scheduler.has_work:
        engine.step()
    return seqs


def stream_generate(engine, prompts: List[List[int]], gen_config=None):
    """Continuous batching with per-token delivery: yields (prompt index, Sequence, new token ids) after every engine step
    in which that prompt produced tokens; the last yield of a prompt has Sequence.finished set (it carries the final tokens,
    or an empty list when the prompt ended without producing any in that step -- cancelled, out of KV memory)."""
    inst = TurboMindInstance(engine)
    cfgs = _per_prompt(gen_config, len(prompts))
    seqs = [_submit_or_reject(inst, engine, p, g) for p, g in zip(prompts, cfgs)]
    sent = [0] * len(seqs)
    done = [False] * len(seqs)
    while not all(done):
        if engine.scheduler.has_work:
            engine.step()
        for i, s in enumerate(seqs):
            if done[i]:
                continue
            if len(s.generated) > sent[i]:
                new = s.generated[sent[i]:]
                sent[i] = len(s.generated)
                done[i] = s.finished                 # the chunk that carries the last token is the final one
                yield i, s, list(new)
            elif s.finished:
                done[i] = True
                yield i, s, []
