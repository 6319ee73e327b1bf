#!/usr/bin/env python
"""Generates tests/golden/ref_attention_*.npz on a GPU box: inputs and outputs of the REFERENCE's own kernels
(oracle/_ref/libtm_ref_attention.so = TurboMind ProcessKV_v2 + dispatchDecoding built in place for sm_100a).
The fixtures let the CPU test-suite (tests/test_golden_cpu.py) pin the numpy oracle to what the reference actually
computes, without a GPU and without /root/reference.

    gpurun -- 'python tools/make_golden.py && cp /tmp/golden/*.npz gpurun_out/'     (then move into tests/golden/)
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lmdeploy_b200 import _lib as L        # noqa: E402  (argument structs only; every kernel below is the reference's)
from oracle import attention as OA         # noqa: E402
from oracle import ref_lib                 # noqa: E402
from oracle.numerics import F16            # noqa: E402


def run_case(kv_bits, rope_name, out_dir):
    ref = ref_lib.load()
    hist = np.array([0, 5, 70, 130], np.int32)
    B, Hq, Hkv, D, layers, layer = len(hist), 8, 2, 128, 2, 1
    rope = {"default": OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e6),
            "llama3": OA.RopeParam(type=OA.ROPE_LLAMA3, dim=128, base=5e5, factor=8.0)}[rope_name]
    layout = OA.KvLayout(Hkv, 64, D, kv_bits)
    bps = [(int(h) + 1 + 63) // 64 for h in hist]
    nblk = sum(bps)
    block_bytes = layers * layout.layer_size()
    rng = np.random.default_rng(100 + kv_bits)
    perm = rng.permutation(nblk)
    tables, o = [], 0
    for n in bps:
        tables.append(perm[o:o + n].tolist())
        o += n
    pool = torch.zeros(nblk, block_bytes, dtype=torch.uint8, device="cuda")
    ptrs = torch.tensor([pool.data_ptr() + int(b) * block_bytes for t in tables for b in t], dtype=torch.int64,
                        device="cuda")
    cu_blocks = torch.tensor(np.concatenate([[0], np.cumsum(bps)]), dtype=torch.int32, device="cuda")

    def base():
        p = L.AttentionParams()
        p.block_ptrs, p.cu_block_nums = ptrs.data_ptr(), cu_blocks.data_ptr()
        p.layer_offset, p.block_len = layer * layout.layer_size(), 64
        p.num_heads, p.num_kv_heads, p.size_per_head = Hq, Hkv, D
        p.inv_sqrt_dh = float(np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D)))))
        p.rope = L.RopeConfig(rope.type, rope.dim, rope.base, rope.factor, rope.max_position_embeddings,
                              rope.llama3_low_freq_factor, rope.llama3_high_freq_factor,
                              rope.llama3_original_max_position_embeddings, rope.yarn_attention_factor,
                              rope.yarn_beta_fast, rope.yarn_beta_slow, 0)
        p.quant_policy, p.dtype = (0 if kv_bits == 16 else kv_bits), 0
        p.stream = torch.cuda.current_stream().cuda_stream
        return p

    # history through the reference's ProcessKV (prefill writer)
    T = int(hist.sum())
    kv_hist = (rng.standard_normal((T, 2 * Hkv * D))).astype(np.float16)
    d_kv = torch.from_numpy(kv_hist).cuda()
    cu_h = torch.tensor(np.concatenate([[0], np.cumsum(hist)]), dtype=torch.int32, device="cuda")
    p = base()
    p.k, p.v, p.stride = d_kv.data_ptr(), d_kv.data_ptr() + Hkv * D * 2, 2 * Hkv * D
    p.cu_q_len = p.cu_k_len = cu_h.data_ptr()
    p.token_num, p.batch_size, p.max_k_len = T, B, int(hist.max())
    a, keep = ref_lib.args_from_b200(p, torch, max_q_len=int(hist.max()))
    assert ref.ref_process_kv(C.byref(a)) == 0
    torch.cuda.synchronize()
    pool_after_prefill = pool.cpu().numpy().copy()

    # one decode step through the reference's dispatchDecoding
    qkv = rng.standard_normal((B, (Hq + 2 * Hkv) * D)).astype(np.float16)
    d_qkv = torch.from_numpy(qkv).cuda()
    out = torch.zeros(B, Hq * D, dtype=torch.float16, device="cuda")
    cu_q = torch.arange(B + 1, dtype=torch.int32, device="cuda")
    cu_k = torch.tensor(np.concatenate([[0], np.cumsum(hist + 1)]), dtype=torch.int32, device="cuda")
    p = base()
    p.out, p.q = out.data_ptr(), d_qkv.data_ptr()
    p.k, p.v = d_qkv.data_ptr() + Hq * D * 2, d_qkv.data_ptr() + (Hq + Hkv) * D * 2
    p.stride = (Hq + 2 * Hkv) * D
    p.cu_q_len, p.cu_k_len = cu_q.data_ptr(), cu_k.data_ptr()
    p.token_num, p.batch_size, p.max_k_len = B, B, int(hist.max()) + 1
    a, keep = ref_lib.args_from_b200(p, torch)
    assert ref.ref_decode_attention(C.byref(a)) == 0
    torch.cuda.synchronize()
    name = f"ref_attention_kv{kv_bits}_{rope_name}.npz"
    np.savez_compressed(os.path.join(out_dir, name), hist=hist, tables=np.array([t + [-1] * (max(bps) - len(t))
                                                                                 for t in tables], np.int32),
                        kv_hist=kv_hist, qkv=qkv, pool_after_prefill=pool_after_prefill,
                        pool_after_decode=pool.cpu().numpy(), out=out.cpu().numpy(),
                        meta=np.array([B, Hq, Hkv, D, layers, layer, kv_bits, rope.type], np.int64),
                        rope_base=np.float64(rope.base), rope_factor=np.float64(rope.factor),
                        source=np.bytes_(ref.ref_version()))
    print("wrote", name)


if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for kv_bits, rope_name in [(8, "default"), (4, "default"), (16, "default"), (8, "llama3")]:
        run_case(kv_bits, rope_name, out_dir)
