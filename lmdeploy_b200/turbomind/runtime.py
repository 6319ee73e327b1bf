"""Device runtime of the decode path: KV block pool, prepared weights and the native decode-step driver.

Mirrors the roles of (reference, for orientation):
  * SequenceManager / BlockManager KV pool   src/turbomind/engine/block.h, memory/
  * LinearWeight::prepare / ModelWeight tree src/turbomind/models/linear_weight.cc:101-324, model_weight.h
  * LanguageModel::Run(kForward)             src/turbomind/models/language_model.cc:493-542
All compute is in libb200tm.so (hand-written sm_100a CUDA); torch is used only for device memory and streams.
"""
import ctypes as C
import os
import math

import torch

from .. import _lib as L
from .specs import ModelSpec, RopeSpec

_TORCH_T = {"f16": torch.float16, "bf16": torch.bfloat16}
_DT = {"f16": 0, "bf16": 1}


def kv_layer_bytes(kv_heads, block_len, head_dim, quant_policy):
    """block::Layout::layer_size (kernels/attention/block.h:172-176)."""
    q_bits = 16 if quant_policy == 0 else quant_policy
    data = block_len * head_dim * q_bits // 8
    param = block_len * 4 if quant_policy else 0
    return kv_heads * 2 * (data + param)


class KvCachePool:
    """Flat pool of fixed-size KV blocks; one block holds block_len tokens of ALL layers
    (unified_attention_layer.cc:163-171)."""

    def __init__(self, num_blocks, layers, kv_heads, head_dim=128, block_len=64, quant_policy=8, device="cuda"):
        self.layer_bytes = kv_layer_bytes(kv_heads, block_len, head_dim, quant_policy)
        self.block_bytes = self.layer_bytes * layers
        self.block_len = block_len
        self.num_blocks = num_blocks
        self.storage = torch.zeros(num_blocks * self.block_bytes, dtype=torch.uint8, device=device)
        self.base = self.storage.data_ptr()
        self.free = list(range(num_blocks - 1, -1, -1))

    def allocate(self, n):
        if len(self.free) < n:
            raise RuntimeError(f"KV pool exhausted: want {n} blocks, {len(self.free)} free")
        return [self.free.pop() for _ in range(n)]

    def release(self, blocks):
        self.free.extend(reversed(blocks))

    def ptr(self, block_id):
        return self.base + block_id * self.block_bytes


class PreparedLinear:
    def __init__(self, lib, K, N, fmt, dtype, device):
        self.desc = L.LinearDesc(K, N, fmt, _DT[dtype], 128, None, None)
        nbytes = lib.b200_linear_prepared_bytes(C.byref(self.desc))
        if nbytes == 0:
            raise L.B200LibraryError(f"linear [{K}x{N}] fmt={fmt}: {lib.b200_last_error().decode()}")
        self.storage = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.desc.prepared = self.storage.data_ptr()
        self.bias = None

    def prepare(self, lib, weight, scales=None, zeros=None, transposed=False, bias=None):
        st = torch.cuda.current_stream().cuda_stream
        L.check(lib.b200_linear_prepare(C.byref(self.desc), weight.data_ptr(),
                                        scales.data_ptr() if scales is not None else None,
                                        zeros.data_ptr() if zeros is not None else None,
                                        1 if transposed else 0, self.storage.data_ptr(), st), "b200_linear_prepare")
        if bias is not None:
            self.bias = bias
            self.desc.bias = bias.data_ptr()


class DecodeModel:
    """Owns the prepared weights of one TP rank and the native b200_model handle."""

    def __init__(self, spec: ModelSpec, *, quant_policy=8, block_len=64, max_batch=64, tp_size=1, tp_rank=0,
                 device=None, rope_exact_math=False):
        """rope_exact_math: correctly rounded RoPE transcendentals (comparable with the CPU oracle) instead of the
        reference release build's hardware approximations (the default; bit-identical K pages to TurboMind)."""
        self.lib = L.load()
        # tuned GEMM launch plans (tools/gemm_sweep.cu writes the lines): B200_GEMM_PLANS=<file>, else the table measured on a
        # B200 that ships with the package (lmdeploy_b200/gemm_plans.txt; B200_GEMM_PLANS=none = heuristic only) -- the
        # counterpart of the reference's GEMM dispatch cache (kernels/gemm/dispatch_cache.cu)
        plans = os.environ.get("B200_GEMM_PLANS")
        if plans is None:
            default = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gemm_plans.txt")
            plans = default if os.path.exists(default) else None
        elif plans in ("", "none", "0"):
            plans = None
        if plans:
            n = self.lib.b200_linear_load_plans(plans.encode())
            if n < 0:
                raise L.B200LibraryError(f"B200_GEMM_PLANS={plans}: {self.lib.b200_last_error().decode()}")
        self.spec = spec
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.tp_size, self.tp_rank = tp_size, tp_rank
        self.quant_policy, self.block_len, self.max_batch = quant_policy, block_len, max_batch
        assert spec.heads % tp_size == 0 and spec.kv_heads % tp_size == 0 and spec.inter % tp_size == 0
        self.heads, self.kv_heads = spec.heads // tp_size, spec.kv_heads // tp_size
        self.inter = spec.inter // tp_size
        self.vocab_pad = (spec.vocab + 128 * tp_size - 1) // (128 * tp_size) * (128 * tp_size)
        self.vocab_local = self.vocab_pad // tp_size
        self.T = _TORCH_T[spec.dtype]
        cfg = L.ModelConfig()
        cfg.num_layers, cfg.hidden = spec.layers, spec.hidden
        cfg.num_heads, cfg.num_kv_heads, cfg.head_dim = self.heads, self.kv_heads, spec.head_dim
        cfg.inter_size, cfg.vocab, cfg.embedding_rows = self.inter, self.vocab_local, spec.vocab
        cfg.norm_eps = spec.norm_eps
        cfg.inv_sqrt_dh = float(torch.tensor(1.4426950408889634, dtype=torch.float32)
                                / torch.sqrt(torch.tensor(float(spec.head_dim), dtype=torch.float32)))
        r = spec.rope or RopeSpec(type=spec.rope_type, dim=spec.head_dim, base=spec.rope_theta, factor=spec.rope_factor)
        cfg.rope = L.RopeConfig(**r.config_fields(), exact_math=1 if rope_exact_math else 0)
        cfg.dtype, cfg.quant_policy, cfg.block_len = _DT[spec.dtype], quant_policy, block_len
        cfg.max_batch, cfg.tp_size, cfg.tp_rank = max_batch, tp_size, tp_rank
        self.cfg = cfg
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.b200_model_create(C.byref(self.handle), C.byref(cfg)), "b200_model_create")
        self._keep = []           # tensors the native model points into
        self.layer_bytes = kv_layer_bytes(self.kv_heads, block_len, spec.head_dim, quant_policy)
        self.weight_bytes = 0     # algorithmic weight bytes streamed per step (this rank)
        self.lm_head_bytes = 0
        self.comm = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _set_linear(self, layer, which, K, N, weight, scales=None, zeros=None, transposed=False, fmt=None):
        # per-linear format: a quantised checkpoint may keep individual linears dense (`modules_to_not_convert`)
        if fmt is None:
            fmt = L.FMT_U4 if scales is not None else L.FMT_DENSE
        if fmt == L.FMT_U4 and (scales is None or zeros is None):
            raise ValueError("a u4 linear needs scales and zero points")
        if fmt == L.FMT_FP8_COL and (scales is None or scales.dtype != torch.float32 or tuple(scales.shape) != (K // 128, N)):
            raise ValueError(f"an fp8 linear [{K}x{N}] needs f32 scales [{K // 128}, {N}]")
        pl = PreparedLinear(self.lib, K, N, fmt, self.spec.dtype, self.device)
        pl.prepare(self.lib, weight, scales, zeros, transposed)
        L.check(self.lib.b200_model_set_linear(self.handle, layer, which, C.byref(pl.desc)), "set_linear")
        self._keep.append(pl)
        nbytes = pl.storage.numel()
        if which == L.W_LM_HEAD:
            self.lm_head_bytes += nbytes
        else:
            self.weight_bytes += nbytes
        return pl

    def _set_tensor(self, layer, which, t):
        t = t.to(self.device, self.T).contiguous()
        self._keep.append(t)
        L.check(self.lib.b200_model_set_tensor(self.handle, layer, which, t.data_ptr()), "set_tensor")

    def load_tm_weights(self, provider):
        """provider(name) -> tensors in the TurboMind pre-prepare layout of THIS rank:
           'tok_embeddings' [vocab, hidden]; 'norm' [hidden]; 'output' [hidden, vocab_local] (dense [K,N]);
           per layer i: f'layers.{i}.attention_norm', f'layers.{i}.ffn_norm',
           f'layers.{i}.{w_qkv|wo|w1w3|w2}' -> dict(weight=..., scales=..., zeros=...) (u4: int32 [K, N/8], T [K/128,N]),
           dict(weight= uint8 e4m3 [K, N], col_scales= f32 [K/128, N]) (fp8) or dict(weight= T [K, N])."""
        s = self.spec
        D = s.head_dim
        with torch.cuda.device(self.device):
            self._set_tensor(0, L.T_EMBEDDING, provider("tok_embeddings"))
            self._set_tensor(0, L.T_FINAL_NORM, provider("norm"))
            shapes = {
                "w_qkv": (s.hidden, (self.heads + 2 * self.kv_heads) * D, L.W_QKV),
                "wo": (self.heads * D, s.hidden, L.W_O),
                "w1w3": (s.hidden, 2 * self.inter, L.W_GATE_UP),
                "w2": (self.inter, s.hidden, L.W_DOWN),
            }
            for i in range(s.layers):
                self._set_tensor(i, L.T_ATTN_NORM, provider(f"layers.{i}.attention_norm"))
                self._set_tensor(i, L.T_FFN_NORM, provider(f"layers.{i}.ffn_norm"))
                if s.qkv_bias:
                    b = provider(f"layers.{i}.qkv_bias")
                    if tuple(b.shape) != ((self.heads + 2 * self.kv_heads) * D,):
                        raise ValueError(f"layers.{i}.qkv_bias has shape {tuple(b.shape)}")
                    self._set_tensor(i, L.T_QKV_BIAS, b)
                if s.qk_norm:
                    for which, nm in ((L.T_Q_NORM, "q_norm"), (L.T_K_NORM, "k_norm")):
                        w = provider(f"layers.{i}.{nm}")
                        if tuple(w.shape) != (D,):
                            raise ValueError(f"layers.{i}.{nm} has shape {tuple(w.shape)}")
                        self._set_tensor(i, which, w)
                for name, (K, N, which) in shapes.items():
                    w = provider(f"layers.{i}.{name}")
                    if "col_scales" in w:       # fp8 e4m3 bytes [K, N] + f32 scales [K/128, N]
                        self._set_linear(i, which, K, N, w["weight"], w["col_scales"], fmt=L.FMT_FP8_COL)
                    else:
                        self._set_linear(i, which, K, N, w["weight"], w.get("scales"), w.get("zeros"))
                    del w
            out = provider("output")
            self._set_linear(0, L.W_LM_HEAD, s.hidden, self.vocab_local, out["weight"],
                             transposed=out.get("transposed", False), fmt=0)
            torch.cuda.synchronize()

    def load_synthetic(self, seed=0):
        """Random weights of the real shapes (no checkpoints exist in the image)."""
        s = self.spec
        g = torch.Generator(device=self.device)
        g.manual_seed(seed * 1000 + self.tp_rank)
        gr = torch.Generator(device=self.device)      # replicated tensors: same on every rank
        gr.manual_seed(seed)
        dev, T = self.device, self.T

        def dense(K, N, gen):
            return (torch.randn(K, N, generator=gen, device=dev, dtype=torch.float32) * (0.1 / math.sqrt(K))).to(T)

        def u4(K, N):
            qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K, N // 8), generator=g, device=dev, dtype=torch.int32)
            sc = ((torch.rand(K // 128, N, generator=g, device=dev) + 0.5) * (0.02 / math.sqrt(K))).to(T)
            zs = torch.randint(6, 11, (K // 128, N), generator=g, device=dev).to(T)
            return dict(weight=qw, scales=sc, zeros=zs)

        def fp8(K, N):
            # the usual fp8 quantiser: scale = amax / 448 per (128-k group, column), weights stored as e4m3(w / scale)
            w = torch.randn(K, N, generator=g, device=dev, dtype=torch.float32) * (0.1 / math.sqrt(K))
            sc = (w.view(K // 128, 128, N).abs().amax(1) / 448.0).clamp_min(1e-12)
            w8 = (w.view(K // 128, 128, N) / sc[:, None]).view(K, N).to(torch.float8_e4m3fn).view(torch.uint8)
            return dict(weight=w8.contiguous(), col_scales=sc.contiguous())

        def provider(name):
            if name == "tok_embeddings":
                return (torch.randn(s.vocab, s.hidden, generator=gr, device=dev) * 0.02).to(T)
            if name.endswith("q_norm") or name.endswith("k_norm"):
                return (1.0 + 0.1 * torch.randn(s.head_dim, generator=gr, device=dev)).to(T)
            if name == "norm" or name.endswith("_norm"):
                return (1.0 + 0.02 * torch.randn(s.hidden, generator=gr, device=dev)).to(T)
            if name.endswith("qkv_bias"):
                return (0.1 * torch.randn((self.heads + 2 * self.kv_heads) * s.head_dim, generator=g, device=dev)).to(T)
            if name == "output":
                return dict(weight=dense(s.hidden, self.vocab_local, g))
            kind = name.split(".")[-1]
            D = s.head_dim
            K, N = {"w_qkv": (s.hidden, (self.heads + 2 * self.kv_heads) * D), "wo": (self.heads * D, s.hidden),
                    "w1w3": (s.hidden, 2 * self.inter), "w2": (self.inter, s.hidden)}[kind]
            if s.weight_format == "awq":
                return u4(K, N)
            if s.weight_format == "fp8":
                return fp8(K, N)
            return dict(weight=dense(K, N, g))

        self.load_tm_weights(provider)

    # ------------------------------------------------------------------ step
    def set_comm(self, comm_handle):
        self.comm = comm_handle
        L.check(self.lib.b200_model_set_comm(self.handle, comm_handle), "set_comm")

    def decode_step(self, *, batch, token_ids, cu_q_len, cu_k_len, block_ptrs, cu_block_nums, out_tokens,
                    max_k_len, finished=None, advance=False, stream=None, token_num=None):
        """All arguments are device tensors (int32 / int64 for block_ptrs / uint8 for finished).
        token_num > batch: a mixed prefill/decode step (b200_model_chunk_step) -- token_ids / out_tokens have token_num
        rows and cu_q_len gives each sequence's rows."""
        si = L.StepInputs()
        si.batch_size = batch
        si.token_ids, si.cu_q_len, si.cu_k_len = token_ids.data_ptr(), cu_q_len.data_ptr(), cu_k_len.data_ptr()
        si.finished = finished.data_ptr() if finished is not None else None
        si.block_ptrs, si.cu_block_nums = block_ptrs.data_ptr(), cu_block_nums.data_ptr()
        si.layer_stride_bytes = self.layer_bytes
        si.max_k_len = max_k_len
        si.out_tokens = out_tokens.data_ptr() if out_tokens is not None else None
        si.advance = 1 if advance else 0
        st = (stream or torch.cuda.current_stream()).cuda_stream
        if token_num is not None and token_num != batch:
            L.check(self.lib.b200_model_chunk_step(self.handle, C.byref(si), int(token_num), st), "b200_model_chunk_step")
            return
        L.check(self.lib.b200_model_decode_step(self.handle, C.byref(si), st), "b200_model_decode_step")

    def logits(self, batch):
        """View of the step's logits [batch, vocab_local] (T)."""
        ptr = self.lib.b200_model_buffer(self.handle, 0)
        return _as_tensor(ptr, (self.max_batch, self.vocab_local), self.T, self.device)[:batch]

    def prefill_kv_random(self, pool: KvCachePool, block_ptrs, cu_block_nums, batch, ctx, seed=1, chunk=None):
        """Fill positions [0, ctx) of every sequence and layer with N(0,1) keys/values through the
        parity-checked b200_process_kv (bench / test setup; realistic quant params)."""
        s = self.spec
        g = torch.Generator(device=self.device)
        g.manual_seed(seed)
        D = s.head_dim
        cu = (torch.arange(batch + 1, dtype=torch.int32, device=self.device) * ctx).contiguous()
        for layer in range(s.layers):
            kv = torch.randn(batch * ctx, 2 * self.kv_heads * D, generator=g, device=self.device,
                             dtype=torch.float32).to(self.T)
            p = L.AttentionParams()
            p.k, p.v, p.stride = kv.data_ptr(), kv.data_ptr() + self.kv_heads * D * 2, 2 * self.kv_heads * D
            p.cu_q_len = p.cu_k_len = cu.data_ptr()
            p.block_ptrs, p.cu_block_nums = block_ptrs.data_ptr(), cu_block_nums.data_ptr()
            p.layer_offset, p.block_len = layer * self.layer_bytes, self.block_len
            p.token_num, p.batch_size, p.max_k_len = batch * ctx, batch, ctx
            p.num_heads, p.num_kv_heads, p.size_per_head = self.heads, self.kv_heads, D
            p.inv_sqrt_dh = self.cfg.inv_sqrt_dh
            p.rope = self.cfg.rope
            p.quant_policy, p.dtype = self.quant_policy, self.cfg.dtype
            p.stream = torch.cuda.current_stream().cuda_stream
            L.check(self.lib.b200_process_kv(C.byref(p)), "b200_process_kv")
            torch.cuda.synchronize()
            del kv

    # algorithmic bytes: weights + lm_head + KV read (+ write) for this rank
    def step_bytes(self, batch, ctx):
        q_bits = 16 if self.quant_policy == 0 else self.quant_policy
        D = self.spec.head_dim
        per_tok = 2 * self.kv_heads * (D * q_bits // 8 + (4 if self.quant_policy else 0))
        kv_read = batch * ctx * self.spec.layers * per_tok
        kv_write = batch * self.spec.layers * per_tok
        return dict(weights=self.weight_bytes, lm_head=self.lm_head_bytes, kv_read=kv_read, kv_write=kv_write,
                    total=self.weight_bytes + self.lm_head_bytes + kv_read + kv_write)


def _as_tensor(ptr, shape, dtype, device):
    """Wrap raw device memory owned by the native library as a torch tensor (no copy)."""
    n = 1
    for d in shape:
        n *= d
    nbytes = n * torch.empty((), dtype=dtype).element_size()

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {
        "shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2,
    }
    t = torch.as_tensor(h, device=device)
    return t.view(dtype).view(*shape)
