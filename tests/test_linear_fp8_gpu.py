"""FP8-weight GEMM (weight_format 3: e4m3 + per-(128-k-group, column) scales, the form the loader emits) vs the numpy oracle.
First device run: round 2 (profiles/r02_first_call.md)."""
import numpy as np
import pytest

from oracle import linear as OL
from oracle.numerics import BF16, F16, rn

from test_linear_gpu import DevLinear, _assert_close

pytestmark = pytest.mark.gpu


def _fp8_cols(rng, K, N):
    """Random weights quantised to e4m3 with one scale per (128-k group, column)."""
    w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
    scale = (np.abs(w.reshape(K // 128, 128, N)).max(axis=1) / 448.0).astype(np.float32)
    w8 = OL.f32_to_e4m3(w / np.repeat(scale, 128, 0))
    return w8, scale


@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("shape", [(1024, 256), (512, 384), (4096, 4096)])
@pytest.mark.parametrize("m", [1, 64])
def test_fp8_per_column_scales(lib, shape, m, dtype):
    """weight_format 3: fp8 e4m3 with scales f32 [K/128, N]; same arithmetic as format 2
    (w = T(e4m3) *_T T(scale), transform.h:92-97)."""
    K, N = shape
    rng = np.random.default_rng(K + N + m)
    w8, scale = _fp8_cols(rng, K, N)
    lin = DevLinear(lib, K, N, 3, dtype, w8, scales=scale)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), dtype)
    got = lin.forward(x)
    ref = OL.linear(x, OL.dequant_fp8_cols(w8, scale, 128, dtype), dtype)
    _assert_close(got, ref, atol=2e-4 if dtype == F16 else 2e-3, rtol=2e-3 if dtype == F16 else 1.6e-2)
    lin_t = DevLinear(lib, K, N, 3, dtype, np.ascontiguousarray(w8.T), scales=scale, transposed=1)
    assert (lin_t.prepared == lin.prepared).all()


def test_fp8_block_scales_expanded_equal_format_2(lib):
    """A checkpoint's 128x128 block scales expanded along N must give bit-identical outputs through format 3."""
    K, N, m = 2048, 1024, 48
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((K, N)) * 0.1 / np.sqrt(K)).astype(np.float32)
    blk = (np.abs(w.reshape(K // 128, 128, N // 128, 128)).max(axis=(1, 3)) / 448.0).astype(np.float32)
    w8 = OL.f32_to_e4m3(w / np.repeat(np.repeat(blk, 128, 0), 128, 1))
    x = rn(rng.standard_normal((m, K)).astype(np.float32), BF16)
    y2 = DevLinear(lib, K, N, 2, BF16, w8, scales=blk).forward(x)
    y3 = DevLinear(lib, K, N, 3, BF16, w8, scales=np.repeat(blk, 128, 1)).forward(x)
    assert np.array_equal(y2, y3)


def test_fp8_gated_silu_on_interleaved_columns(lib):
    """Fused gate|up with element interleave -- impossible with 128-wide N blocks, the reason format 3 exists."""
    K, I, m = 2048, 1024, 32
    rng = np.random.default_rng(6)
    g8, gs = _fp8_cols(rng, K, I)
    u8, us = _fp8_cols(rng, K, I)
    w8 = np.stack([g8, u8], -1).reshape(K, 2 * I)
    sc = np.stack([gs, us], -1).reshape(K // 128, 2 * I)
    lin = DevLinear(lib, K, 2 * I, 3, BF16, w8, scales=sc)
    x = rn(rng.standard_normal((m, K)).astype(np.float32), BF16)
    got = lin.forward(x, epilogue=1)
    ref = OL.linear_gated_silu(x, OL.dequant_fp8_cols(w8, sc, 128, BF16), BF16)
    _assert_close(got, ref, atol=2e-3, rtol=1.6e-2)
