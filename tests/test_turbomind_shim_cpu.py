"""CPU checks of the `_turbomind` shim (lmdeploy_b200/shim/_turbomind.py): the module surface the reference's linear test
harness touches (tests/turbomind/linear/linear.py) and the torch-side quantisers, which need no GPU.  The GEMM calls
themselves run on a B200 (tests/test_reference_linear_harness_gpu.py)."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tm():
    spec = importlib.util.spec_from_file_location("_turbomind", os.path.join(ROOT, "lmdeploy_b200", "shim", "_turbomind.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_surface_used_by_the_reference_harness(tm):
    for name in ("DataType", "Epilogue", "Tensor", "LinearConfig", "LinearWeight", "LlamaLinear", "ResolveLinearWeightFormat",
                 "QuantizeSymmBlock", "DequantizeSymmBlock", "QuantizeSymm", "DequantizeSymm", "QuantizeGroupwise",
                 "MakeStridedPtrs", "invokeMoeDispatch", "invokeMoeCombine", "from_dlpack", "from_dlpack_with_strides",
                 "generic_copy_on_stream", "create_device_context"):
        assert hasattr(tm, name), name
    for name in ("TYPE_FP16", "TYPE_BF16", "TYPE_UINT4", "TYPE_UINT8", "TYPE_FP8_E4M3", "TYPE_FP32"):
        assert hasattr(tm.DataType, name), name
    assert hasattr(tm.Epilogue, "kNone") and hasattr(tm.Epilogue, "kGatedSilu")
    cfg = tm.LinearConfig()
    cfg.input_dim, cfg.output_dim, cfg.data_type = 256, 512, tm.DataType.TYPE_FP16
    cfg.format = tm.ResolveLinearWeightFormat(tm.DataType.TYPE_FP16, tm.DataType.TYPE_UINT4, 128, 1)
    assert cfg.format.dtype == tm.DataType.TYPE_UINT4 and cfg.format.block_in == 128


@pytest.mark.parametrize("shape", [(256, 512), (256, 256), (384, 256)])     # (K, N); K == N / 2 once scrambled the nibbles
def test_quantize_groupwise_packs_k_major_views(tm, shape):
    K, N = shape
    g = 128
    torch.manual_seed(K + N)
    w = (torch.randn(K, N) * 0.05).half()
    quant = torch.zeros(K, N // 2, dtype=torch.uint8)
    scales = torch.zeros(K // g, N, dtype=torch.float16)
    zeros = torch.zeros(K // g, N, dtype=torch.float16)
    deq = torch.zeros(K, N, dtype=torch.float16)
    T = tm.Tensor
    tm.QuantizeGroupwise(quant=T(quant, tm.DataType.TYPE_UINT4).t(), scales=T(scales).t(), zeros=T(zeros).t(),
                         dequant=T(deq).t(), src=T(w).t(), group_size=g)
    q = torch.empty(K, N, dtype=torch.float32)
    q[:, 0::2] = (quant & 15).float()
    q[:, 1::2] = (quant >> 4).float()
    s = scales.float().repeat_interleave(g, dim=0)
    z = zeros.float().repeat_interleave(g, dim=0)
    assert q.min() >= 0 and q.max() <= 15 and z.min() >= 0 and z.max() <= 15
    torch.testing.assert_close(((q - z) * s).half(), deq, rtol=0, atol=0)
    # 4-bit round-to-nearest: within half a step of the source (a full step where the integer zero-point clamps)
    assert ((deq.float() - w.float()).abs() <= s * 1.01).all()


def test_symm_block_fp8_round_trip(tm):
    if not hasattr(torch, "float8_e4m3fn"):
        pytest.skip("no fp8 dtype in this torch")
    torch.manual_seed(3)
    K, N = 256, 384
    w = (torch.randn(K, N) * 0.05).bfloat16()
    out, scale = tm.QuantizeSymmBlock(src=tm.Tensor(w))
    assert out.shape == [K, N] and scale.shape == [K // 128, N // 128]
    back = tm.DequantizeSymmBlock(src=out, scale=scale)
    err = (back._t.float() - w.float()).abs().max().item()
    assert err <= 0.07 * w.float().abs().max().item()      # e4m3: 3 mantissa bits
