"""Checkpoint-format un-packing (lmdeploy_b200/turbomind/checkpoint_formats.py) against outputs of the reference's own
lmdeploy/turbomind/weight_format.py, captured by tools/make_loader_golden.py into tests/golden/ref_weight_formats.npz.
Integer/byte work: every comparison is exact."""
import os

import numpy as np
import pytest
import torch

from lmdeploy_b200.turbomind import checkpoint_formats as CF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_weight_formats.npz"))


def tensors_of(prefix):
    head = prefix + ".in"
    return {k[len(head):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(head)}


@pytest.mark.parametrize("prefix,fmt", [("awq", CF.AwqFormat()), ("gptq", CF.GptqFormat()),
                                        ("ct", CF.CompressedTensorsFormat())])
def test_int4_families_match_reference_normalize_and_pack(prefix, fmt):
    av = tensors_of(prefix)
    assert fmt.accepts(av)
    lin = fmt.read(av)
    assert lin.kind == "u4"
    assert np.array_equal(lin.q.numpy(), GOLD[f"{prefix}.tm.weight"])
    assert np.array_equal(lin.scales.to(torch.float16).numpy(), GOLD[f"{prefix}.tm.scales"])
    assert np.array_equal(lin.zeros.to(torch.float16).numpy(), GOLD[f"{prefix}.tm.zeros"])
    assert np.array_equal(CF.pack_u4_row(lin.q.contiguous()).numpy(), GOLD[f"{prefix}.packed"])
    if f"{prefix}.dequant" in GOLD.files:
        assert np.array_equal(lin.dequant().numpy(), GOLD[f"{prefix}.dequant"])


@pytest.mark.parametrize("prefix,fmt,drop", [("gptq", CF.GptqFormat(), ".qzeros"),
                                             ("ct", CF.CompressedTensorsFormat(), ".weight_zero_point")])
def test_missing_zero_points_are_synthesised_like_the_reference(prefix, fmt, drop):
    av = tensors_of(prefix)
    del av[drop]
    lin = fmt.read(av)
    assert np.array_equal(lin.zeros.to(torch.float16).numpy(), GOLD[f"{prefix}.synth_zeros"])
    assert (lin.zeros == CF.SYMMETRIC_ZERO).all()


def test_fp8_block_scaled_orientation():
    av = tensors_of("fp8")
    fmt = CF.Fp8Format()
    assert fmt.accepts(av)
    lin = fmt.read(av)
    assert lin.kind == "fp8"
    assert np.array_equal(lin.w8.numpy(), GOLD["fp8.tm.weight"])
    blocks = GOLD["fp8.tm.scales"]                                       # reference: [K/128, N/128]
    assert tuple(lin.scales.shape) == (blocks.shape[0], lin.w8.shape[1])
    assert np.array_equal(lin.scales.numpy(), np.repeat(blocks, 128, 1))   # expanded along N, nothing else
    # the float8 view of the same bytes is accepted as well
    av[".weight"] = av[".weight"].view(torch.float8_e4m3fn)
    assert np.array_equal(fmt.read(av).w8.numpy(), GOLD["fp8.tm.weight"])
    # with per-column scales, column operations keep weight and scale together
    idx = torch.tensor([300, 5, 129])
    c = lin.cols(idx)
    assert torch.equal(c.dequant(), lin.dequant()[:, idx])
    r = lin.rows(128, 256)
    assert torch.equal(r.dequant(), lin.dequant()[128:256])
    with pytest.raises(ValueError, match="splits a quantisation group"):
        lin.rows(0, 64)
    assert torch.equal(CF.Lin.cat([c, c]).dequant(), torch.cat([c.dequant(), c.dequant()], 1))


def test_accepts_truth_table_matches_reference():
    g = torch.Generator().manual_seed(0)
    K, N, G = 256, 32, 128
    i32 = lambda *s: torch.randint(-2 ** 31, 2 ** 31 - 1, s, generator=g, dtype=torch.int32)
    f16 = lambda *s: torch.rand(s, generator=g).to(torch.float16)
    cases = {
        "awq": {".qweight": i32(K, N // 8), ".scales": f16(K // G, N), ".qzeros": i32(K // G, N // 8)},
        "gptq": {".qweight": i32(K // 8, N), ".scales": f16(K // G, N), ".qzeros": i32(K // G, N // 8)},
        "ct": {".weight_packed": i32(N, K // 8), ".weight_scale": f16(N, K // G)},
        "fp8": {".weight": torch.zeros(N, K, dtype=torch.uint8), ".weight_scale_inv": torch.ones(1, 2)},
        "dense": {".weight": torch.zeros(N, K, dtype=torch.float16)},
    }
    ours = {"awq": CF.AwqFormat(), "gptq": CF.GptqFormat(), "compressed-tensors": CF.CompressedTensorsFormat(),
            "fp8": CF.Fp8Format(), "trivial": CF.DenseFormat()}
    assert list(GOLD["accepts.cases"]) == list(cases)
    table = GOLD["accepts.table"]
    for r, case in enumerate(cases.values()):
        for c, name in enumerate(GOLD["accepts.formats"]):
            assert bool(ours[str(name)].accepts(case)) == bool(table[r, c]), (r, name)


def test_resolver_falls_through_to_dense_and_fails_loudly():
    av = tensors_of("awq")
    tensors = {"a.b" + k: v for k, v in av.items()}
    tensors["head.weight"] = torch.zeros(8, 16, dtype=torch.float16)
    rd = CF.FormatResolver("awq")
    assert rd.read(tensors, "a.b").kind == "u4"
    dense = rd.read(tensors, "head")                      # un-converted module inside a quantised checkpoint
    assert dense.kind == "dense" and tuple(dense.w.shape) == (16, 8)
    with pytest.raises(KeyError, match="no checkpoint tensors"):
        rd.read(tensors, "missing")
    tensors["odd.weight"] = torch.zeros(8, 16, dtype=torch.int8)
    with pytest.raises(ValueError, match="no weight format accepts"):
        rd.read(tensors, "odd")
    with pytest.raises(ValueError, match="unknown checkpoint quantisation"):
        CF.FormatResolver("int3")
    # a GPTQ resolver must not swallow AWQ-shaped tensors (and vice versa)
    with pytest.raises(ValueError):
        CF.FormatResolver("gptq").read(tensors, "a.b")


def test_gptq_act_order_is_rejected():
    av = tensors_of("gptq")
    K = av[".qweight"].shape[0] * 8
    av[".g_idx"] = torch.arange(K, dtype=torch.int32) // 128
    CF.GptqFormat().read(av)
    av[".g_idx"] = torch.flip(av[".g_idx"], [0])
    with pytest.raises(NotImplementedError, match="act-order"):
        CF.GptqFormat().read(av)


def test_group_size_is_checked_against_the_tensors():
    av = tensors_of("awq")
    with pytest.raises(ValueError, match="not group 64"):
        CF.AwqFormat(group=64).read(av)


def test_lin_slicing_keeps_groups_whole_and_biases_aligned():
    av = tensors_of("awq")
    av[".bias"] = torch.arange(32, dtype=torch.float16)
    lin = CF.AwqFormat().read(av)
    idx = torch.tensor([3, 1, 2])
    c = lin.cols(idx)
    assert torch.equal(c.q, lin.q[:, idx]) and torch.equal(c.zeros, lin.zeros[:, idx])
    assert torch.equal(c.bias, torch.tensor([3.0, 1.0, 2.0]))
    r = lin.rows(128, 256)
    assert r.q.shape[0] == 128 and r.scales.shape[0] == 1 and torch.equal(r.scales[0], lin.scales[1])
    with pytest.raises(ValueError, match="splits a quantisation group"):
        lin.rows(64, 256)
    cat = CF.Lin.cat([c, c])
    assert cat.q.shape[1] == 6 and cat.bias.shape[0] == 6
    with pytest.raises(NotImplementedError, match="different formats"):
        CF.Lin.cat([c, CF.Lin(w=torch.zeros(256, 3))])


QC = CF.parse_quantization_config


def test_quantization_config_admission():
    assert QC(None) == (None, 0) and QC({}) == (None, 0)
    assert QC({"quant_method": "awq", "bits": 4, "group_size": 128, "version": "gemm"}) == ("awq", 128)
    assert QC({"quant_method": "awq", "w_bit": 4, "q_group_size": 128, "version": "GEMM"}) == ("awq", 128)
    assert QC({"quant_method": "gptq", "bits": 4, "group_size": 128, "sym": True, "desc_act": False}) == ("gptq", 128)
    ct = {"quant_method": "compressed-tensors",
          "config_groups": {"group_0": {"format": "pack-quantized",
                                        "weights": {"num_bits": 4, "type": "int", "group_size": 128}}}}
    assert QC(ct) == ("compressed-tensors", 128)
    assert QC({"quant_method": "fp8", "weight_block_size": [128, 128]}) == ("fp8", 128)
    bad = [
        {"quant_method": "awq", "bits": 4, "group_size": 128, "version": "gemv"},
        {"quant_method": "awq", "bits": 4, "group_size": 64, "version": "gemm"},
        {"quant_method": "awq", "bits": 8, "group_size": 128, "version": "gemm"},
        {"quant_method": "gptq", "bits": 4, "group_size": 128, "desc_act": True},
        {"quant_method": "gptq", "bits": 4, "group_size": 128, "sym": False},
        {"quant_method": "fp8", "weight_block_size": [64, 64]},
        {"quant_method": "bitsandbytes"},
        {"quant_method": "compressed-tensors",
         "config_groups": {"group_0": {"format": "int-quantized",
                                       "weights": {"num_bits": 8, "type": "int", "group_size": 128}}}},
        {"quant_method": "compressed-tensors",
         "config_groups": {"group_0": {"format": "pack-quantized",
                                       "weights": {"num_bits": 4, "type": "int", "group_size": 32}}}},
    ]
    for q in bad:
        with pytest.raises(NotImplementedError):
            QC(q)
