"""GPU parity: b200_decode_attention (C ABI) vs the CPU oracle on identical seeded inputs.

Bars (DESIGN.md numerics contract):
  * KV-cache bytes written by the decode prologue (quantised k/v of the new token, their (scale, zero))
    -- BIT-EXACT vs the oracle.
  * attention output (fp16/bf16): |diff| <= 2e-3 abs for N(0,1) inputs (fp32 accumulation order and
    ex2.approx differ from the oracle's float64 accumulation; the reference's own PyTorch-engine kernel
    test uses atol 1e-3 on the same quantity, tests/pytorch/kernel/test_paged_attention.py:495-511).
"""
import numpy as np
import pytest

from oracle import attention as OA
from oracle.numerics import F16, BF16

from harness import PagedCase

pytestmark = pytest.mark.gpu

HIST = [0, 1, 15, 63, 64, 65, 127, 130, 200, 517]


def _check(case, out, pool, ref, ref_pool, tol, finished=None):
    if finished is not None:
        keep = ~np.asarray(finished, bool)
        out, ref = out[keep], ref[keep]
    mism = np.nonzero(pool != ref_pool)
    assert mism[0].size == 0, f"KV cache bytes differ at {mism[0].size} positions; first block {mism[0][:4]} byte {mism[1][:4]}"
    err = np.abs(out - ref)
    assert np.isfinite(out).all()
    assert err.max() <= tol, f"max abs err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
@pytest.mark.parametrize("heads", [(8, 2), (8, 1), (12, 2)])
def test_decode_attention_f16(kv_bits, heads):
    case = PagedCase(batch_hist=HIST, num_heads=heads[0], num_kv_heads=heads[1], kv_bits=kv_bits, seed=kv_bits)
    out, pool = case.run_device()
    ref = case.run_oracle()
    _check(case, out, pool, ref, case.cache.pool, 2e-3)


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
def test_decode_attention_bf16(kv_bits):
    case = PagedCase(batch_hist=HIST, num_heads=8, num_kv_heads=2, kv_bits=kv_bits, dtype=BF16, seed=7)
    out, pool = case.run_device()
    ref = case.run_oracle()
    _check(case, out, pool, ref, case.cache.pool, 1.6e-2)   # bf16 ulp at |x|~1 is 2^-8


@pytest.mark.parametrize("rope", ["llama3", "linear", "null", "partial", "yarn"])
def test_rope_variants(rope):
    r = {
        "llama3": OA.RopeParam(type=OA.ROPE_LLAMA3, dim=128, base=5e5, factor=8.0),
        "linear": OA.RopeParam(type=OA.ROPE_LINEAR, dim=128, base=1e4, factor=2.0),
        "null": OA.RopeParam(type=OA.ROPE_NULL, dim=128),
        "partial": OA.RopeParam(type=OA.ROPE_DEFAULT, dim=64, base=1e4),
        "yarn": OA.RopeParam(type=OA.ROPE_YARN, dim=128, base=1e4, factor=4.0, max_position_embeddings=4096,
                             yarn_attention_factor=1.1),
    }[rope]
    case = PagedCase(batch_hist=[0, 70, 300], num_heads=8, num_kv_heads=2, kv_bits=8, rope=r, seed=11)
    out, pool = case.run_device()
    ref = case.run_oracle()
    _check(case, out, pool, ref, case.cache.pool, 2e-3)


def test_bias_blocklen_finished():
    fin = [0, 1, 0, 0, 1, 0]
    case = PagedCase(batch_hist=[5, 100, 64, 191, 30, 256], num_heads=16, num_kv_heads=2, kv_bits=8, block_len=128,
                     bias=True, seed=5, layers=3, layer=2)
    out, pool = case.run_device(finished=fin)
    ref = case.run_oracle(finished=fin)
    _check(case, out, pool, ref, case.cache.pool, 2e-3, finished=fin)


def test_long_context_split_across_ctas():
    """One long sequence: its tiles are spread over every CTA (stream-K) and merged by the fix-up kernel."""
    case = PagedCase(batch_hist=[9000, 3], num_heads=8, num_kv_heads=1, kv_bits=8, seed=2)
    out, pool = case.run_device()
    ref = case.run_oracle()
    _check(case, out, pool, ref, case.cache.pool, 2e-3)


def test_run_to_run_determinism():
    case = PagedCase(batch_hist=[700, 33, 1500, 64], num_heads=8, num_kv_heads=2, kv_bits=4, seed=3)
    out1, _ = case.run_device()
    out2, _ = case.run_device()
    assert (out1 == out2).all()


def test_full_size_sampled():
    """BASELINE config 2 shape (bs=64, ctx 2048, 32 q heads / 8 kv heads, KV int8): cache written by
    b200_process_kv on the GPU, attention checked against the oracle on sampled (sequence, kv-head) pairs
    that read back the GPU-written pages; plus size-independent properties."""
    import ctypes as C
    import torch
    from lmdeploy_b200 import _lib as L
    from harness import to_dev_T, from_dev_T, rand_T
    lib = L.load()
    B, ctx, Hq, Hkv, D, layers, layer = 64, 2048, 32, 8, 128, 2, 1
    rng = np.random.default_rng(0)
    layout = OA.KvLayout(Hkv, 64, D, 8)
    blocks_per_seq = (ctx + 1 + 63) // 64
    nblk = B * blocks_per_seq
    block_bytes = layers * layout.layer_size()
    pool = torch.zeros(nblk, block_bytes, dtype=torch.uint8, device="cuda")
    perm = rng.permutation(nblk)
    tables = perm.reshape(B, blocks_per_seq)
    ptrs = torch.tensor((pool.data_ptr() + tables.reshape(-1).astype(np.int64) * block_bytes), dtype=torch.int64,
                        device="cuda")
    cu_blocks = torch.arange(B + 1, dtype=torch.int32, device="cuda") * blocks_per_seq
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e6)

    def base_params():
        p = L.AttentionParams()
        p.block_ptrs, p.cu_block_nums = ptrs.data_ptr(), cu_blocks.data_ptr()
        p.layer_offset, p.block_len = layer * layout.layer_size(), 64
        p.num_heads, p.num_kv_heads, p.size_per_head = Hq, Hkv, D
        p.inv_sqrt_dh = float(np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D)))))
        p.rope = L.RopeConfig(rope.type, rope.dim, rope.base, 0.0, 0, 1.0, 4.0, 8192, 1.0, 32.0, 1.0, 1)  # exact_math
        p.quant_policy, p.dtype = 8, 0
        p.stream = torch.cuda.current_stream().cuda_stream
        return p

    # prefill the cache: ctx tokens per sequence through b200_process_kv
    kv = torch.randn(B * ctx, 2 * Hkv * D, device="cuda", dtype=torch.float16)
    p = base_params()
    p.k, p.v, p.stride = kv.data_ptr(), kv.data_ptr() + Hkv * D * 2, 2 * Hkv * D
    cu_q = torch.arange(B + 1, dtype=torch.int32, device="cuda") * ctx
    p.cu_q_len = p.cu_k_len = cu_q.data_ptr()
    p.token_num, p.batch_size, p.max_k_len = B * ctx, B, ctx
    L.check(lib.b200_process_kv(C.byref(p)), "b200_process_kv")

    # decode step
    qkv = torch.randn(B, (Hq + 2 * Hkv) * D, device="cuda", dtype=torch.float16)
    out = torch.zeros(B, Hq * D, device="cuda", dtype=torch.float16)
    p = base_params()
    p.out, p.q = out.data_ptr(), qkv.data_ptr()
    p.k, p.v, p.stride = qkv.data_ptr() + Hq * D * 2, qkv.data_ptr() + (Hq + Hkv) * D * 2, (Hq + 2 * Hkv) * D
    cu_ql = torch.arange(B + 1, dtype=torch.int32, device="cuda")
    cu_kl = torch.arange(B + 1, dtype=torch.int32, device="cuda") * (ctx + 1)
    p.cu_q_len, p.cu_k_len = cu_ql.data_ptr(), cu_kl.data_ptr()
    p.token_num, p.batch_size, p.max_k_len = B, B, ctx + 1
    ws = torch.zeros(lib.b200_decode_attention_workspace_bytes(B, Hq, Hkv), dtype=torch.uint8, device="cuda")
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    L.check(lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")
    torch.cuda.synchronize()
    out_h = from_dev_T(out, F16).reshape(B, Hq, D)
    assert np.isfinite(out_h).all()

    # oracle on sampled (b, h): read back GPU-written pages
    cache = OA.PagedKvCache(nblk, layers, layout, F16)
    cache.pool = pool.cpu().numpy()
    qkv_h = from_dev_T(qkv, F16)
    hist = np.full(B, ctx, np.int32)
    q, k, v = OA.decode_prologue(qkv_h, None, None, None, hist, Hq, Hkv, D, rope, F16)
    G = Hq // Hkv
    worst = 0.0
    for b, h in [(0, 0), (17, 3), (63, 7), (31, 5)]:
        K, V = cache.read_tokens(tables[b], layer, h, ctx + 1)
        # the token appended by the decode prologue must equal the oracle's quantisation of (k, v)
        ref_cache = OA.PagedKvCache(1, 1, layout, F16)
        ref_cache.write_token(0, 0, h, 0, k[b, h], v[b, h])
        Kr, Vr = ref_cache.read_tokens([0], 0, h, 1)
        assert (K[ctx] == Kr[0]).all() and (V[ctx] == Vr[0]).all()
        for g in range(G):
            o, _, _, _ = OA.softmax_pv(q[b, h * G + g], K, V, p.inv_sqrt_dh, F16)
            worst = max(worst, float(np.abs(o - out_h[b, h * G + g]).max()))
    assert worst <= 2e-3, worst
    # property: output is a convex combination of V rows -> bounded by per-dim min/max of V
    K, V = cache.read_tokens(tables[5], layer, 2, ctx + 1)
    o = out_h[5, 2 * G:(2 + 1) * G]
    assert (o <= V.max(0) + 2e-3).all() and (o >= V.min(0) - 2e-3).all()


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
@pytest.mark.parametrize("dtype", [F16, BF16])
@pytest.mark.parametrize("flags", [8])
def test_fused_kernel_equals_the_three_kernel_form(kv_bits, dtype, flags):
    """The single-launch form (flags 8: in-register RoPE of q, the owning warp quantises + stores the new token's K/V and
    patches its tile, the last-arriving CTA merges straddling segments) against the production form (prologue kernel, main
    kernel, fix-up kernel; debug_flags bit 2): KV pages and outputs
    BIT-IDENTICAL -- same arithmetic, same merge order.  Histories cover an empty cache, block / tile boundaries and
    sequences long enough to straddle CTAs (the grid has more CTAs than there are tiles here: every multi-tile sequence is
    spread over CTAs with empty ones in between); bias + llama3 rope on."""
    rope = OA.RopeParam(type=OA.ROPE_LLAMA3, dim=128, base=5e5, factor=8.0, llama3_low_freq_factor=1.0,
                        llama3_high_freq_factor=4.0, llama3_original_max_position_embeddings=8192)
    hist = [0, 1, 63, 64, 65, 127, 128, 700, 2047, 4100]
    # exact_math=False: the reference release build's transcendentals, the arithmetic the engine runs with
    a = PagedCase(batch_hist=hist, num_heads=8, num_kv_heads=2, kv_bits=kv_bits, dtype=dtype, seed=5, bias=True, rope=rope,
                  exact_math=False)
    b = PagedCase(batch_hist=hist, num_heads=8, num_kv_heads=2, kv_bits=kv_bits, dtype=dtype, seed=5, bias=True, rope=rope,
                  exact_math=False)
    out_f, pool_f = a.run_device(debug_flags=flags)
    out_u, pool_u = b.run_device(debug_flags=4)
    assert (pool_f == pool_u).all(), "KV pages differ between the fused and the three-kernel form"
    assert (out_f.view(np.uint16) == out_u.view(np.uint16)).all(), "outputs differ between the fused and the three-kernel form"
    # and the workspace's arrival counters are back at zero: a second call on the same workspace gives the same result
    from lmdeploy_b200 import _lib as L
    import ctypes as C
    import torch
    lib = L.load()
    p = a.params(lib, workspace=a.d_ws)
    p.debug_flags = flags
    L.check(lib.b200_decode_attention(C.byref(p)), "b200_decode_attention")
    torch.cuda.synchronize()
    from harness import from_dev_T
    assert (from_dev_T(a.d_out, dtype).view(np.uint16) == out_f.view(np.uint16)).all()


@pytest.mark.parametrize("flags", [0, 4, 8])
def test_partials_of_an_earlier_call_do_not_leak_into_a_sparse_one(flags):
    """A workspace that served a big call (every CTA of the grid left fp32 partials) then serves a small one whose tiles are
    fewer than the grid's CTAs: the pieces of a multi-tile sequence then sit in CTAs with EMPTY CTAs in between, whose
    partial slots still hold the big call's values.  The merge must visit the pieces only (round 1 walked the whole CTA
    range and was right only on a zeroed workspace)."""
    import ctypes as C
    import torch
    from lmdeploy_b200 import _lib as L
    from harness import from_dev_T
    lib = L.load()
    big = PagedCase(batch_hist=[3000] * 24, num_heads=8, num_kv_heads=2, kv_bits=8, seed=21)
    big.upload()
    pb = big.params(lib)
    pb.debug_flags = flags
    L.check(lib.b200_decode_attention(C.byref(pb)), "b200_decode_attention")
    torch.cuda.synchronize()
    small = PagedCase(batch_hist=[130, 517, 70, 200], num_heads=8, num_kv_heads=2, kv_bits=8, seed=22)
    small.upload()
    ps = small.params(lib, workspace=big.d_ws)             # the used workspace, not a zeroed one
    ps.debug_flags = flags
    L.check(lib.b200_decode_attention(C.byref(ps)), "b200_decode_attention")
    torch.cuda.synchronize()
    out = from_dev_T(small.d_out, F16)
    ref = small.run_oracle()
    assert np.isfinite(out).all() and np.abs(out - ref).max() <= 2e-3, float(np.abs(out - ref).max())


@pytest.mark.parametrize("kv_bits", [8, 4, 16])
def test_early_kv_streaming_with_a_caller_plan_is_bit_identical(kv_bits):
    """With the caller's plan the producer warp starts copying KV tiles before griddepcontrol.wait and waits only before the
    first tile that holds a new token (written by the prologue kernel); outputs and pages must equal the plain call's bit for
    bit -- histories chosen so that CTAs start on first, middle and last tiles, and on sequences of one tile."""
    hist = [0, 1, 63, 64, 65, 127, 128, 700, 2047, 2048, 3000, 5]
    a = PagedCase(batch_hist=hist, num_heads=8, num_kv_heads=2, kv_bits=kv_bits, seed=21)
    out_a, pool_a = a.run_device()
    b = PagedCase(batch_hist=hist, num_heads=8, num_kv_heads=2, kv_bits=kv_bits, seed=21)
    for _ in range(3):                       # repeated: a race with the prologue would not show every time
        out_b, pool_b = b.run_device(external_plan=True)
        assert (pool_a == pool_b).all()
        assert (out_a == out_b).all()
