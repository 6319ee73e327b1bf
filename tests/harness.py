"""Shared test harness: builds paged KV caches with the CPU oracle, uploads them, and drives the C ABI
(libb200tm.so) through ctypes with torch tensors as device memory.  GPU tests only."""
import ctypes as C

import numpy as np

from oracle import attention as OA
from oracle.numerics import F16, BF16, to_bits, from_bits, rn


def torch_mod():
    import torch
    return torch


def np_T(dtype):
    return np.float16 if dtype == F16 else None


def to_dev_T(x_f32, dtype, device="cuda"):
    """float32 numpy (exact in T) -> torch tensor of dtype T on device."""
    torch = torch_mod()
    bits = to_bits(x_f32, dtype).astype(np.int16)
    t = torch.from_numpy(bits).to(device)
    return t.view(torch.float16 if dtype == F16 else torch.bfloat16)


def from_dev_T(t, dtype):
    torch = torch_mod()
    bits = t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
    return from_bits(bits, dtype)


def rand_T(rng, shape, dtype, scale=1.0):
    return rn((rng.standard_normal(shape) * scale).astype(np.float32), dtype)


class PagedCase:
    """A batch of sequences with a paged KV cache filled by the oracle."""

    def __init__(self, *, batch_hist, num_heads, num_kv_heads, kv_bits, dtype=F16, block_len=64, layers=2,
                 layer=1, rope=None, seed=0, bias=False, extra_blocks=3, exact_math=True):
        # exact_math: correctly rounded RoPE transcendentals on the device (comparable with the numpy oracle);
        # False = the reference release build's arithmetic (compared against oracle/_ref, the reference kernels)
        self.exact_math = exact_math
        self.rng = np.random.default_rng(seed)
        self.hist = np.asarray(batch_hist, np.int32)
        self.B = len(batch_hist)
        self.Hq, self.Hkv, self.D = num_heads, num_kv_heads, 128
        self.dtype = dtype
        self.kv_bits = kv_bits
        self.layer = layer
        self.rope = rope if rope is not None else OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=10000.0)
        self.layout = OA.KvLayout(head_num=num_kv_heads, block_len=block_len, head_dim=128, q_bits=kv_bits)
        blocks_per_seq = [(int(h) + 1 + block_len - 1) // block_len for h in self.hist]
        nblk = sum(blocks_per_seq) + extra_blocks
        self.cache = OA.PagedKvCache(nblk, layers, self.layout, dtype)
        perm = self.rng.permutation(nblk)
        self.block_tables = []
        o = 0
        for n in blocks_per_seq:
            self.block_tables.append(perm[o:o + n].tolist())
            o += n
        self.cu_block_nums = np.concatenate([[0], np.cumsum(blocks_per_seq)]).astype(np.int32)
        self.bias = None
        if bias:
            self.bias = (rand_T(self.rng, (num_heads * 128,), dtype, 0.1),
                         rand_T(self.rng, (num_kv_heads * 128,), dtype, 0.1),
                         rand_T(self.rng, (num_kv_heads * 128,), dtype, 0.1))
        # history K/V (pre-rotation values), written through the oracle's ProcessKV restatement
        for b in range(self.B):
            h = int(self.hist[b])
            if h == 0:
                continue
            k = rand_T(self.rng, (h, num_kv_heads, 128), dtype)
            v = rand_T(self.rng, (h, num_kv_heads, 128), dtype)
            OA.process_kv(self.cache, self.block_tables[b], layer, 0, k, v, self.rope,
                          None if not bias else self.bias[1], None if not bias else self.bias[2])
        self.qkv = rand_T(self.rng, (self.B, (num_heads + 2 * num_kv_heads) * 128), dtype)

    # ---- oracle ----
    def run_oracle(self, finished=None):
        qb = kb = vb = None
        if self.bias:
            qb, kb, vb = self.bias
        return OA.decode_attention(self.qkv, self.cache, self.block_tables, self.hist, self.layer, self.Hq,
                                   self.Hkv, self.rope, q_bias=qb, k_bias=kb, v_bias=vb, finished=finished)

    # ---- device ----
    def upload(self):
        torch = torch_mod()
        self.d_pool = torch.from_numpy(self.cache.pool.copy()).cuda()
        base = self.d_pool.data_ptr()
        ptrs = []
        for bt in self.block_tables:
            ptrs += [base + int(b) * self.cache.block_bytes for b in bt]
        self.d_block_ptrs = torch.tensor(ptrs, dtype=torch.int64, device="cuda")
        self.d_cu_block_nums = torch.from_numpy(self.cu_block_nums).cuda()
        self.d_qkv = to_dev_T(self.qkv, self.dtype)
        k_len = self.hist + 1
        self.d_cu_k_len = torch.from_numpy(np.concatenate([[0], np.cumsum(k_len)]).astype(np.int32)).cuda()
        self.d_cu_q_len = torch.arange(self.B + 1, dtype=torch.int32, device="cuda")
        self.d_out = torch.zeros(self.B, self.Hq * 128, dtype=self.d_qkv.dtype, device="cuda")
        if self.bias:
            self.d_bias = [to_dev_T(x, self.dtype) for x in self.bias]

    def params(self, lib, finished=None, workspace=None):
        from lmdeploy_b200 import _lib as L
        torch = torch_mod()
        p = L.AttentionParams()
        D = 128
        qkv_ptr = self.d_qkv.data_ptr()
        p.out = self.d_out.data_ptr()
        p.q = qkv_ptr
        p.k = qkv_ptr + self.Hq * D * 2
        p.v = qkv_ptr + (self.Hq + self.Hkv) * D * 2
        p.stride = (self.Hq + 2 * self.Hkv) * D
        if self.bias:
            p.q_bias, p.k_bias, p.v_bias = [t.data_ptr() for t in self.d_bias]
        p.cu_q_len = self.d_cu_q_len.data_ptr()
        p.cu_k_len = self.d_cu_k_len.data_ptr()
        if finished is not None:
            self.d_finished = torch.from_numpy(np.asarray(finished, np.uint8)).cuda()
            p.finished = self.d_finished.data_ptr()
        p.block_ptrs = self.d_block_ptrs.data_ptr()
        p.cu_block_nums = self.d_cu_block_nums.data_ptr()
        p.layer_offset = self.cache.layer_offset(self.layer)
        p.block_len = self.layout.block_len
        p.token_num = self.B
        p.batch_size = self.B
        p.max_k_len = int(self.hist.max()) + 1
        p.num_heads, p.num_kv_heads, p.size_per_head = self.Hq, self.Hkv, D
        p.inv_sqrt_dh = float(np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D)))))
        p.window_size = 0
        r = self.rope
        p.rope = L.RopeConfig(r.type, r.dim, r.base, r.factor, r.max_position_embeddings,
                              r.llama3_low_freq_factor, r.llama3_high_freq_factor,
                              r.llama3_original_max_position_embeddings, r.yarn_attention_factor,
                              r.yarn_beta_fast, r.yarn_beta_slow, 1 if self.exact_math else 0)
        p.quant_policy = 0 if self.kv_bits == 16 else self.kv_bits
        p.dtype = 0 if self.dtype == F16 else 1
        nbytes = lib.b200_decode_attention_workspace_bytes(self.B, self.Hq, self.Hkv)
        self.d_ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda") if workspace is None else workspace
        p.workspace = self.d_ws.data_ptr()
        p.workspace_bytes = self.d_ws.numel()
        p.stream = torch.cuda.current_stream().cuda_stream
        return p

    def run_device(self, finished=None, debug_flags=0, external_plan=False):
        from lmdeploy_b200 import _lib as L
        torch = torch_mod()
        lib = L.load()
        self.upload()
        p = self.params(lib, finished)
        p.debug_flags = debug_flags
        if external_plan:
            # the caller's plan (as b200_model_decode_step passes one per step): lets the main kernel's producer warp stream
            # KV tiles before griddepcontrol.wait (KernelParams::early)
            self.d_plan = torch.zeros(self.B + 1, dtype=torch.int32, device="cuda")
            L.check(lib.b200_decode_attention_plan(self.d_plan.data_ptr(), p.cu_k_len, p.finished, self.B, p.stream), "plan")
            p.plan = self.d_plan.data_ptr()
        L.check(lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")
        torch.cuda.synchronize()
        return from_dev_T(self.d_out, self.dtype), self.d_pool.cpu().numpy()
