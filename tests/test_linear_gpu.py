"""GPU parity: b200_linear_* (tcgen05 W4A16 / dense GEMM) vs the CPU oracle.

Bars: the dequantised weight is bit-identical by construction (f16 fma); the product is accumulated in fp32
on the tensor core in an unspecified order, so outputs are compared at 1 fp16 ulp of the output scale:
|diff| <= 2e-4 + 2e-3*|ref|  (reference gate for the same quantity: max_abs <= 0.25,
tests/turbomind/linear/fixture.py:34-42; we are ~1000x tighter)."""
import ctypes as C

import numpy as np
import pytest

from oracle import linear as OL
from oracle.numerics import F16, BF16, rn

from harness import to_dev_T, from_dev_T

pytestmark = pytest.mark.gpu


def _mk_u4(rng, K, N):
    w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
    q, s, z = OL.quantize_groupwise_u4(w, 128)
    return q, s, z


class DevLinear:
    def __init__(self, lib, K, N, fmt, dtype, weight, scales=None, zeros=None, bias=None, transposed=0):
        import torch
        from lmdeploy_b200 import _lib as L
        self.lib, self.L, self.torch = lib, L, torch
        self.d = L.LinearDesc(K, N, fmt, 0 if dtype == F16 else 1, 128, None, None)
        nbytes = lib.b200_linear_prepared_bytes(C.byref(self.d))
        assert nbytes > 0, lib.b200_last_error()
        self.prepared = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        self.d.prepared = self.prepared.data_ptr()
        if bias is not None:
            self.bias = to_dev_T(bias, dtype)
            self.d.bias = self.bias.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        if fmt == 1:
            qw = torch.from_numpy(OL.pack_u4_row(weight)).cuda()
            ds, dz = to_dev_T(scales, dtype), to_dev_T(zeros, dtype)
            L.check(lib.b200_linear_prepare(C.byref(self.d), qw.data_ptr(), ds.data_ptr(), dz.data_ptr(), 0,
                                            self.prepared.data_ptr(), st), "prepare")
        elif fmt in (2, 3):
            dw = torch.from_numpy(np.ascontiguousarray(weight)).cuda()           # e4m3 bits, uint8
            ds = torch.from_numpy(np.ascontiguousarray(scales, np.float32)).cuda()
            L.check(lib.b200_linear_prepare(C.byref(self.d), dw.data_ptr(), ds.data_ptr(), None, transposed,
                                            self.prepared.data_ptr(), st), "prepare")
        else:
            dw = to_dev_T(weight, dtype)
            L.check(lib.b200_linear_prepare(C.byref(self.d), dw.data_ptr(), None, None, transposed,
                                            self.prepared.data_ptr(), st), "prepare")
        torch.cuda.synchronize()
        self.dtype = dtype
        self.K, self.N = K, N
        self.ws = torch.zeros(lib.b200_linear_workspace_bytes(C.byref(self.d), 128), dtype=torch.uint8, device="cuda")

    def forward(self, x, epilogue=0):
        torch, L, lib = self.torch, self.L, self.lib
        m = x.shape[0]
        dx = to_dev_T(x, self.dtype)
        nout = self.N // 2 if epilogue == 1 else self.N
        y = torch.zeros(m, nout, dtype=dx.dtype, device="cuda")
        L.check(lib.b200_linear_forward(C.byref(self.d), dx.data_ptr(), self.K, m, y.data_ptr(), nout, epilogue,
                                        self.ws.data_ptr(), self.ws.numel(), torch.cuda.current_stream().cuda_stream),
                "b200_linear_forward")
        torch.cuda.synchronize()
        return from_dev_T(y, self.dtype)


def _assert_close(got, ref, atol=2e-4, rtol=2e-3):
    err = np.abs(got - ref)
    lim = atol + rtol * np.abs(ref)
    assert np.isfinite(got).all()
    bad = err > lim
    assert not bad.any(), f"{bad.sum()} / {bad.size} out of tolerance; max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)} ref {ref.flat[err.argmax()]} got {got.flat[err.argmax()]}"


@pytest.mark.parametrize("shape", [(256, 128), (1024, 256), (4096, 6144), (4096, 4096), (14336, 4096), (2048, 4096), (3072, 6144)])
@pytest.mark.parametrize("m", [1, 16, 64])
def test_w4a16_f16(lib, shape, m):
    K, N = shape
    rng = np.random.default_rng(K + N + m)
    q, s, z = _mk_u4(rng, K, N)
    lin = DevLinear(lib, K, N, 1, F16, q, s, z)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), F16)
    got = lin.forward(x)
    ref = OL.linear(x, OL.dequant_u4(q, s, z, F16), F16)
    _assert_close(got, ref)


def _layout_flag(lin):
    """trailer word 0 of a prepared u4 stream: 1 = the f16 "subnormal" layout (gemm.cuh: kU4FastShift), 0 = plain"""
    import torch
    return int(lin.prepared[-256:-252].view(torch.int32).item())


def test_w4a16_layout_flag_and_out_of_range_scales(lib):
    """f16 streams whose scales all fit use the subnormal layout (nibble at bits 7..10 of an f16 lane = q * 2^-17, scales stored
    times 2^17: the same fma, bit for bit); one scale >= 0.5 keeps the whole tensor in the plain layout; bf16 is always plain.
    Both layouts must give the oracle's numbers."""
    K, N = 1024, 256
    rng = np.random.default_rng(77)
    w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
    x = rn(rng.standard_normal((17, K)).astype(np.float32), F16)
    q, s, z = OL.quantize_groupwise_u4(w, 128)
    lin = DevLinear(lib, K, N, 1, F16, q, s, z)
    assert _layout_flag(lin) == 1
    _assert_close(lin.forward(x), OL.linear(x, OL.dequant_u4(q, s, z, F16), F16))
    assert _layout_flag(DevLinear(lib, K, N, 1, BF16, q, s, z)) == 0
    w2 = w.copy()
    w2[:128, :64] *= 3000.0                       # one 128 x 64 patch with a range of ~60: scales ~4
    q2, s2, z2 = OL.quantize_groupwise_u4(w2, 128)
    assert float(np.abs(s2).max()) > 0.5
    lin2 = DevLinear(lib, K, N, 1, F16, q2, s2, z2)
    assert _layout_flag(lin2) == 0
    _assert_close(lin2.forward(x), OL.linear(x, OL.dequant_u4(q2, s2, z2, F16), F16), atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("m", [3, 33, 128])
def test_w4a16_ragged_m_and_bias(lib, m):
    K, N = 512, 384
    rng = np.random.default_rng(m)
    q, s, z = _mk_u4(rng, K, N)
    bias = rn((rng.standard_normal(N) * 0.05).astype(np.float32), F16)
    lin = DevLinear(lib, K, N, 1, F16, q, s, z, bias=bias)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), F16)
    got = lin.forward(x)
    ref = OL.linear(x, OL.dequant_u4(q, s, z, F16), F16, bias=bias)
    _assert_close(got, ref)


def test_w4a16_gated_silu(lib):
    K, N = 4096, 2048     # N interleaved (gate, up) -> 1024 outputs
    rng = np.random.default_rng(9)
    q, s, z = _mk_u4(rng, K, N)
    lin = DevLinear(lib, K, N, 1, F16, q, s, z)
    x = rn(rng.standard_normal((64, K)).astype(np.float32), F16)
    got = lin.forward(x, epilogue=1)
    ref = OL.linear_gated_silu(x, OL.dequant_u4(q, s, z, F16), F16)
    _assert_close(got, ref, atol=3e-4)


@pytest.mark.parametrize("dtype", [F16, BF16])
@pytest.mark.parametrize("transposed", [0, 1])
def test_dense(lib, dtype, transposed):
    K, N, m = 1024, 512, 40
    rng = np.random.default_rng(3)
    w = rn((rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32), dtype)
    lin = DevLinear(lib, K, N, 0, dtype, w.T.copy() if transposed else w, transposed=transposed)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), dtype)
    got = lin.forward(x)
    ref = OL.linear(x, w, dtype)
    _assert_close(got, ref, atol=2e-4 if dtype == F16 else 2e-3, rtol=2e-3 if dtype == F16 else 1.6e-2)


def test_w4a16_bf16(lib):
    K, N, m = 1024, 256, 16
    rng = np.random.default_rng(4)
    w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
    q, s, z = OL.quantize_groupwise_u4(w, 128)
    s = rn(s, BF16)
    lin = DevLinear(lib, K, N, 1, BF16, q, s, z)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), BF16)
    got = lin.forward(x)
    ref = OL.linear(x, OL.dequant_u4(q, s, z, BF16), BF16)
    _assert_close(got, ref, atol=2e-3, rtol=1.6e-2)


def test_linearity_property_full_size(lib):
    """Size-independent property at the 7B w1w3 shape: f(a*x) == a*f(x) for a power of two (exact in fp
    arithmetic), and f(0) == 0."""
    K, N = 4096, 28672
    rng = np.random.default_rng(5)
    q, s, z = _mk_u4(rng, K, N)
    lin = DevLinear(lib, K, N, 1, F16, q, s, z)
    x = rn(rng.standard_normal((64, K)).astype(np.float32), F16)
    y1 = lin.forward(x)
    y2 = lin.forward(rn(x * 0.5, F16))
    assert (y2 == rn(y1 * 0.5, F16)).mean() > 0.999   # halving is exact except f16 subnormal outputs
    assert (lin.forward(np.zeros_like(x)) == 0).all()
    # spot-check 64 random columns against the oracle
    cols = rng.choice(N, 64, replace=False)
    wd = OL.dequant_u4(q[:, cols], s[:, cols], z[:, cols], F16)
    _assert_close(y1[:, cols], OL.linear(x, wd, F16))


@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("shape", [(1024, 256), (512, 384), (4096, 4096)])
@pytest.mark.parametrize("m", [1, 64])
def test_fp8_block_scaled(lib, shape, m, dtype):
    """FP8 (e4m3, 128x128 f32 block scales) weight-only linear, the non-sm90 path of the reference:
    w = T(e4m3) *_T T(scale_block) (transform.h:92-97, linear_weight.cc:138-150), T x T -> fp32 accumulate."""
    K, N = shape
    rng = np.random.default_rng(K + N + m)
    w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
    # per-block scale = amax / 448 (the usual fp8 block quantiser), weights stored as e4m3 of w / scale
    wb = w.reshape(K // 128, 128, N // 128, 128)
    scale = (np.abs(wb).max(axis=(1, 3)) / 448.0).astype(np.float32)
    w8 = OL.f32_to_e4m3(w / np.repeat(np.repeat(scale, 128, 0), 128, 1))
    lin = DevLinear(lib, K, N, 2, dtype, w8, scales=scale)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), dtype)
    got = lin.forward(x)
    ref = OL.linear(x, OL.dequant_fp8(w8, scale, 128, dtype), dtype)
    _assert_close(got, ref, atol=2e-4 if dtype == F16 else 2e-3, rtol=2e-3 if dtype == F16 else 1.6e-2)
    # the transposed ([N, K]) loader layout must prepare to the same stream
    lin_t = DevLinear(lib, K, N, 2, dtype, np.ascontiguousarray(w8.T), scales=scale, transposed=1)
    assert (lin_t.prepared == lin.prepared).all()


@pytest.mark.parametrize("m", [64, 17])
@pytest.mark.parametrize("shape", [(4096, 4096), (14336, 4096), (2048, 4096)])
def test_fused_residual_rmsnorm_epilogue(lib, shape, m):
    """wo / w2 GEMM with the residual add + RMSNorm fused into the split-K epilogue
    (b200_linear_forward_packed_residual_rmsnorm) vs linear -> residual_rmsnorm of the oracle."""
    import torch
    from lmdeploy_b200 import _lib as L
    K, N = shape
    rng = np.random.default_rng(K + m)
    q, s, z = _mk_u4(rng, K, N)
    lin = DevLinear(lib, K, N, 1, F16, q, s, z)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), F16)
    residual = rn(rng.standard_normal((m, N)).astype(np.float32), F16)
    nw = rn((1 + 0.1 * rng.standard_normal(N)).astype(np.float32), F16)
    mt = (m + 15) // 16 * 16
    dx = to_dev_T(x, F16)
    xp = torch.zeros(K // 8, mt, 8, dtype=torch.float16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200_pack_activations(dx.data_ptr(), K, m, K, mt, 0, xp.data_ptr(), st), "pack")
    d_res = to_dev_T(residual, F16)
    d_nw = to_dev_T(nw, F16)
    out = torch.zeros(m, N, dtype=torch.float16, device="cuda")
    rc = lib.b200_linear_forward_packed_residual_rmsnorm(C.byref(lin.d), xp.data_ptr(), mt, m, d_res.data_ptr(),
                                                         d_nw.data_ptr(), C.c_float(1e-5), out.data_ptr(), 0,
                                                         lin.ws.data_ptr(), lin.ws.numel(), st)
    L.check(rc, "b200_linear_forward_packed_residual_rmsnorm")
    torch.cuda.synchronize()
    y = OL.linear(x, OL.dequant_u4(q, s, z, F16), F16)
    ref_out, ref_res = OL.residual_rmsnorm(y, residual, nw, 1e-5, F16)
    got_res, got_out = from_dev_T(d_res, F16), from_dev_T(out, F16)
    # the GEMM output may differ from the oracle by one fp16 ulp (fp32 summation order), which moves the residual by
    # at most one ulp of its own magnitude
    _assert_close(got_res, ref_res, atol=2e-3, rtol=2e-3)
    _assert_close(got_out, ref_out, atol=4e-3, rtol=4e-3)
    # second call on the same workspace (self-resetting counters)
    d_res2 = to_dev_T(residual, F16)
    L.check(lib.b200_linear_forward_packed_residual_rmsnorm(C.byref(lin.d), xp.data_ptr(), mt, m, d_res2.data_ptr(),
                                                            d_nw.data_ptr(), C.c_float(1e-5), out.data_ptr(), 0,
                                                            lin.ws.data_ptr(), lin.ws.numel(), st), "second call")
    torch.cuda.synchronize()
    assert (from_dev_T(d_res2, F16) == got_res).all() and (from_dev_T(out, F16) == got_out).all()
