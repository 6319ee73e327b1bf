"""Batch runners on top of the native decode step.

StaticBatchRunner drives a fixed batch of sequences (the benchmark / steady-state decode case): it owns the KV
block pool, the block tables and the per-step device bookkeeping, and can replay the whole step as a CUDA graph.
The continuous-batching engine (lmdeploy_b200.turbomind.engine) is built from the same pieces.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from .turbomind.runtime import DecodeModel, KvCachePool


class StaticBatchRunner:
    def __init__(self, model: DecodeModel, batch, max_len, seed=0):
        self.model, self.batch, self.max_len = model, batch, max_len
        dev = model.device
        spec = model.spec
        self.blocks_per_seq = (max_len + model.block_len - 1) // model.block_len
        nblk = batch * self.blocks_per_seq
        self.pool = KvCachePool(nblk, spec.layers, model.kv_heads, spec.head_dim, model.block_len, model.quant_policy,
                                device=dev)
        # random block placement exercises paging (kernels/attention/test_attention.cu:119-129 does the same)
        perm = np.random.default_rng(seed).permutation(nblk).reshape(batch, self.blocks_per_seq)
        if os.environ.get("B200_BENCH_SEQUENTIAL_PAGES"):      # diagnostic: how much does the shuffled placement cost HBM?
            perm = np.arange(nblk).reshape(batch, self.blocks_per_seq)
        self.tables = perm
        self.block_ptrs = torch.tensor(self.pool.base + perm.reshape(-1).astype(np.int64) * self.pool.block_bytes,
                                       dtype=torch.int64, device=dev)
        self.cu_blocks = (torch.arange(batch + 1, dtype=torch.int32, device=dev) * self.blocks_per_seq).contiguous()
        self.cu_q = torch.arange(batch + 1, dtype=torch.int32, device=dev)
        self.cu_k = torch.zeros(batch + 1, dtype=torch.int32, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self.token_ids = torch.randint(0, spec.vocab, (batch,), generator=g, device=dev, dtype=torch.int32)
        self.out_tokens = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.h_in = torch.zeros(batch, dtype=torch.int32).pin_memory()
        self.h_out = torch.zeros(batch, dtype=torch.int32).pin_memory()
        self.cur_len = 0
        # kernels per step: embedding, rmsnorm, plan, rope table, L x (qkv, prologue, attention, fix-up, wo, norm, w13,
        # w2, norm), lm_head, argmax stage 1 + 2, advance  (B200_FUSED_NORM=1 folds the two norms into wo / w2: 7 per layer)
        self.launches_per_step = 4 + (7 if (model.tp_size == 1 and os.environ.get("B200_FUSED_NORM")) else 9) * spec.layers + 4
        self._graph = None

    def prefill_random(self, ctx):
        self.model.prefill_kv_random(self.pool, self.block_ptrs, self.cu_blocks, self.batch, ctx)
        self.reset_lengths(ctx)

    def reset_lengths(self, ctx):
        """Every sequence has `ctx` cached tokens; the next step appends token ctx (k_len = ctx + 1)."""
        self.cu_k.copy_(torch.arange(self.batch + 1, dtype=torch.int32, device=self.cu_k.device) * (ctx + 1))
        self.cur_len = ctx + 1

    def _step(self, advance=True):
        assert self.cur_len <= self.max_len, "runner out of KV capacity"
        self.model.decode_step(batch=self.batch, token_ids=self.token_ids, cu_q_len=self.cu_q, cu_k_len=self.cu_k,
                               block_ptrs=self.block_ptrs, cu_block_nums=self.cu_blocks, out_tokens=self.out_tokens,
                               max_k_len=self.cur_len, advance=advance)
        if advance:
            self.cur_len += 1

    def make_device_step(self, graph=True):
        """Returns a callable running one decode step entirely from device-resident state."""
        if not graph:
            return self._step
        self._step()                      # warm-up outside capture (lazy attribute / occupancy queries)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.model.decode_step(batch=self.batch, token_ids=self.token_ids, cu_q_len=self.cu_q,
                                   cu_k_len=self.cu_k, block_ptrs=self.block_ptrs, cu_block_nums=self.cu_blocks,
                                   out_tokens=self.out_tokens, max_k_len=self.max_len, advance=True)
        self._graph = g

        def replay():
            assert self.cur_len <= self.max_len, "runner out of KV capacity"
            g.replay()
            self.cur_len += 1
        return replay

    def step_host(self, host_tokens):
        """One step through host buffers: H2D token ids (pinned), decode, D2H sampled ids.  Replays the captured
        step when make_device_step(graph=True) built one (the graph reads token_ids / writes out_tokens in place)."""
        self.h_in.copy_(torch.from_numpy(host_tokens))
        self.token_ids.copy_(self.h_in, non_blocking=True)
        if self._graph is not None:
            assert self.cur_len <= self.max_len, "runner out of KV capacity"
            self._graph.replay()
            self.cur_len += 1
        else:
            self._step(advance=True)
        self.h_out.copy_(self.out_tokens, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_out.numpy().copy()

    def time_attention_kernel(self, k_len, iters=20, layer=0, full=False):
        """CUDA-event time of the decode attention at context k_len: decode_attention_kernel alone (prologue / fix-up
        skipped, the default) or -- full=True -- the whole b200_decode_attention call (prologue + main kernel + fix-up), which
        is what compares with the reference's dispatchDecoding (in-kernel prologue + its reduce).
        Algorithmic bytes per launch = batch * k_len * 2 * Hkv * (D * kvbits / 8 + 4)."""
        m = self.model
        spec = m.spec
        D = spec.head_dim
        dev = m.device
        cu_k = (torch.arange(self.batch + 1, dtype=torch.int32, device=dev) * k_len).contiguous()
        qkv = torch.randn(self.batch, (m.heads + 2 * m.kv_heads) * D, device=dev, dtype=torch.float32).to(m.T)
        out = torch.zeros(self.batch, m.heads * D, device=dev, dtype=m.T)
        ws = torch.zeros(m.lib.b200_decode_attention_workspace_bytes(self.batch, m.heads, m.kv_heads),
                         dtype=torch.uint8, device=dev)
        p = L.AttentionParams()
        p.out, p.q = out.data_ptr(), qkv.data_ptr()
        p.k = qkv.data_ptr() + m.heads * D * 2
        p.v = qkv.data_ptr() + (m.heads + m.kv_heads) * D * 2
        p.stride = (m.heads + 2 * m.kv_heads) * D
        p.cu_q_len, p.cu_k_len = self.cu_q.data_ptr(), cu_k.data_ptr()
        p.block_ptrs, p.cu_block_nums = self.block_ptrs.data_ptr(), self.cu_blocks.data_ptr()
        p.layer_offset, p.block_len = layer * m.layer_bytes, m.block_len
        p.token_num, p.batch_size, p.max_k_len = self.batch, self.batch, k_len
        p.num_heads, p.num_kv_heads, p.size_per_head = m.heads, m.kv_heads, D
        p.inv_sqrt_dh, p.rope = m.cfg.inv_sqrt_dh, m.cfg.rope
        p.quant_policy, p.dtype = m.quant_policy, m.cfg.dtype
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
        p.stream = torch.cuda.current_stream().cuda_stream
        # main-kernel-only timing uses the round-1 three-kernel form (debug bit 2) with prologue and fix-up skipped (bits 0, 1):
        # one unfused full call first leaves the rotated q it reads; full=True times the op as callers get it (one fused kernel)
        p.debug_flags = 0 if full else 4
        if full:
            # as a decode step calls the operator: with the step's tile plan computed once up front (b200_model_decode_step),
            # which also lets the main kernel's producer warp stream KV tiles while the prologue kernel runs
            plan = torch.zeros(self.batch + 1, dtype=torch.int32, device=dev)
            L.check(m.lib.b200_decode_attention_plan(plan.data_ptr(), cu_k.data_ptr(), None, self.batch, p.stream), "plan")
            p.plan = plan.data_ptr()
        L.check(m.lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")     # full call once
        p.debug_flags = 0 if full else 7
        # rotate over layers so successive launches never hit L2-resident pages (each layer is >> L2 anyway)
        layers = spec.layers
        for i in range(3):
            p.layer_offset = ((layer + i) % layers) * m.layer_bytes
            L.check(m.lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            p.layer_offset = ((layer + i) % layers) * m.layer_bytes
            L.check(m.lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        q_bits = 16 if m.quant_policy == 0 else m.quant_policy
        nbytes = self.batch * k_len * 2 * m.kv_heads * (D * q_bits // 8 + (4 if m.quant_policy else 0))
        self._attn_params = (p, qkv, out, ws, cu_k, plan if full else None)   # kept alive: bench.py times the reference kernel on the same buffers
        return dict(us=us, bytes=nbytes, gbs=nbytes / (us * 1e-6) / 1e9)
