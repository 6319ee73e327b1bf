"""GPU parity against the REFERENCE ITSELF: the unmodified TurboMind decode-attention / ProcessKV kernels
(oracle/_ref/libtm_ref_attention.so, built in place from /root/reference by oracle/ref_build/Makefile) run on the
same device buffers as ours.  This is what pins the numpy oracle and our kernels to the reference (the
reference stores no golden vectors for this path).

Bars (rope.exact_math = 0, i.e. the arithmetic of the reference's release build: ex2.approx / sin.approx / cos.approx
and the mul+fma contraction its SASS shows):
  * every byte the prologue / ProcessKV writes to the KV pages (K, V, scale, zero): BIT-IDENTICAL;
  * attention output: |ours - reference| <= 2e-3 (fp16) / 1.6e-2 (bf16) on identical cache contents (split counts
    and the fp32 summation order differ by design).
"""
import ctypes as C

import numpy as np
import pytest

from oracle import attention as OA
from oracle import ref_lib
from oracle.numerics import F16, BF16

from harness import PagedCase, from_dev_T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libtm_ref_attention.so not built (make -C oracle/ref_build needs /root/reference)")
    return ref_lib.load()


def _run_both(case, lib, ref):
    import torch
    from lmdeploy_b200 import _lib as L
    # ours
    case.upload()
    p = case.params(lib)
    L.check(lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")
    torch.cuda.synchronize()
    out_ours, pool_ours = from_dev_T(case.d_out, case.dtype), case.d_pool.cpu().numpy()
    # reference, on a fresh copy of the same cache
    case.upload()
    p = case.params(lib)
    a, keep = ref_lib.args_from_b200(p, torch)
    rc = ref.ref_decode_attention(C.byref(a))
    torch.cuda.synchronize()
    assert rc == 0
    out_ref, pool_ref = from_dev_T(case.d_out, case.dtype), case.d_pool.cpu().numpy()
    return out_ours, pool_ours, out_ref, pool_ref


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
@pytest.mark.parametrize("heads", [(32, 8), (8, 8), (16, 2)])
def test_decode_matches_reference_kernel_f16(lib, ref, kv_bits, heads):
    hq, hkv = heads
    case = PagedCase(batch_hist=[0, 1, 63, 64, 65, 200, 777, 1500], num_heads=hq, num_kv_heads=hkv, kv_bits=kv_bits,
                     seed=kv_bits + hq, exact_math=False)
    out_ours, pool_ours, out_ref, pool_ref = _run_both(case, lib, ref)
    assert (pool_ours == pool_ref).all(), int((pool_ours != pool_ref).sum())
    assert np.isfinite(out_ours).all() and np.isfinite(out_ref).all()
    err = np.abs(out_ours - out_ref).max()
    assert err <= 2e-3, err


@pytest.mark.parametrize("kv_bits", [8, 16])
def test_decode_matches_reference_kernel_bf16(lib, ref, kv_bits):
    case = PagedCase(batch_hist=[5, 130, 900], num_heads=32, num_kv_heads=8, kv_bits=kv_bits, dtype=BF16, seed=3,
                     exact_math=False)
    out_ours, pool_ours, out_ref, pool_ref = _run_both(case, lib, ref)
    assert (pool_ours == pool_ref).all(), int((pool_ours != pool_ref).sum())
    assert np.abs(out_ours - out_ref).max() <= 1.6e-2


@pytest.mark.parametrize("rope", ["default", "llama3", "null"])
def test_process_kv_matches_reference_kernel(lib, ref, rope):
    """Prefill-side KV writer: the same 3 x 300 tokens through b200_process_kv and the reference's
    invokeProcessKV_v2 into separate pools."""
    import torch
    from lmdeploy_b200 import _lib as L
    B, T, Hkv, D, layers, layer = 3, 300, 8, 128, 2, 1
    rp = {"default": OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e6),
          "llama3": OA.RopeParam(type=OA.ROPE_LLAMA3, dim=128, base=5e5, factor=8.0, llama3_low_freq_factor=1.0,
                                 llama3_high_freq_factor=4.0, llama3_original_max_position_embeddings=8192),
          "null": OA.RopeParam(type=OA.ROPE_NULL, dim=128, base=1e4)}[rope]
    for kv_bits in (8, 4):
        layout = OA.KvLayout(Hkv, 64, D, kv_bits)
        bps = (T + 63) // 64
        nblk = B * bps
        block_bytes = layers * layout.layer_size()
        g = torch.Generator(device="cuda").manual_seed(kv_bits)
        kv = torch.randn(B * T, 2 * Hkv * D, device="cuda", dtype=torch.float16, generator=g)
        cu = torch.arange(B + 1, dtype=torch.int32, device="cuda") * T
        cu_blocks = torch.arange(B + 1, dtype=torch.int32, device="cuda") * bps
        pools = []
        for which in ("ours", "ref"):
            pool = torch.zeros(nblk, block_bytes, dtype=torch.uint8, device="cuda")
            ptrs = (pool.data_ptr() + torch.arange(nblk, dtype=torch.int64) * block_bytes).cuda()
            p = L.AttentionParams()
            p.k, p.v, p.stride = kv.data_ptr(), kv.data_ptr() + Hkv * D * 2, 2 * Hkv * D
            p.cu_q_len = p.cu_k_len = cu.data_ptr()
            p.block_ptrs, p.cu_block_nums = ptrs.data_ptr(), cu_blocks.data_ptr()
            p.layer_offset, p.block_len = layer * layout.layer_size(), 64
            p.token_num, p.batch_size, p.max_k_len = B * T, B, T
            p.num_heads, p.num_kv_heads, p.size_per_head = Hkv, Hkv, D
            p.inv_sqrt_dh = 0.1275
            p.rope = L.RopeConfig(rp.type, rp.dim, rp.base, rp.factor, rp.max_position_embeddings,
                                  rp.llama3_low_freq_factor, rp.llama3_high_freq_factor,
                                  rp.llama3_original_max_position_embeddings, rp.yarn_attention_factor,
                                  rp.yarn_beta_fast, rp.yarn_beta_slow, 0)
            p.quant_policy, p.dtype = kv_bits, 0
            p.stream = torch.cuda.current_stream().cuda_stream
            if which == "ours":
                L.check(lib.b200_process_kv(C.byref(p)), "b200_process_kv")
            else:
                a, keep = ref_lib.args_from_b200(p, torch, max_q_len=T)
                assert ref.ref_process_kv(C.byref(a)) == 0
            torch.cuda.synchronize()
            pools.append(pool.cpu().numpy().reshape(nblk, layers, -1)[:, layer])
        ours, refp = pools
        per_head = 64 * D * kv_bits // 8
        data = Hkv * 2 * per_head
        # [head][K|V][token][row]: V data of every head must be bit-identical, so must all (scale, zero) of V
        for h in range(Hkv):
            v_lo = (2 * h + 1) * per_head
            assert (ours[:, v_lo:v_lo + per_head] == refp[:, v_lo:v_lo + per_head]).all()
            pv = data + (2 * h + 1) * 64 * 4
            assert (ours[:, pv:pv + 256] == refp[:, pv:pv + 256]).all()
        assert (ours == refp).all(), (rope, kv_bits, int((ours != refp).sum()))
