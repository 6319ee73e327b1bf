"""Generate tests/golden/ref_weight_formats.npz: the reference's checkpoint-format normalisation
(lmdeploy/turbomind/weight_format.py -- AWQ / GPTQ / compressed-tensors / FP8 `normalize`, `pack`, `synthesize_zeros`,
`dequant`) run on small random checkpoint tensors.  Runs in the build container only (needs /root/reference); the
fixture it writes is what travels.

The reference module imports the compiled `_turbomind` extension for its DataType enum only; a stand-in module with
that enum is installed first, and the two pure-Python files are loaded by path so the package `__init__` (which needs
the compiled engine) is not executed.  No reference code is copied: the functions are called, their outputs saved.

  python tools/make_loader_golden.py
"""
import enum
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/lmdeploy/turbomind"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "ref_weight_formats.npz")


def load_reference():
    tm = types.ModuleType("_turbomind")

    class DataType(enum.Enum):
        TYPE_UINT4 = 1
        TYPE_FP8_E4M3 = 2
        TYPE_FP4_E2M1 = 3
        TYPE_FP16 = 4
        TYPE_BF16 = 5

    tm.DataType = DataType
    tm.DataFormat = object
    tm.ResolveLinearWeightFormat = lambda *a: a
    sys.modules["_turbomind"] = tm
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [REF]
    sys.modules["refpkg"] = pkg
    mods = {}
    for name in ("linear", "weight_format"):
        spec = importlib.util.spec_from_file_location(f"refpkg.{name}", os.path.join(REF, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refpkg.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["weight_format"]


def main():
    wf = load_reference()
    g = torch.Generator().manual_seed(20260922)
    K, N, G = 256, 32, 128
    out = {}

    def rnd_i32(*shape):
        return torch.randint(-2 ** 31, 2 ** 31 - 1, shape, generator=g, dtype=torch.int32)

    def rnd_f16(*shape):
        return (torch.rand(shape, generator=g) * 0.02 + 0.001).to(torch.float16)

    def save(prefix, fmt, raw, kinds):
        for suffix, t in raw.items():
            out[f"{prefix}.in{suffix}"] = t.numpy()
        tm = {}
        for suffix, kind in kinds.items():
            if suffix in raw:
                tm[kind] = fmt.normalize(raw[suffix], kind)
                out[f"{prefix}.tm.{kind}"] = tm[kind].contiguous().numpy()
        packed = fmt.pack(tm["weight"].contiguous(), "weight")
        out[f"{prefix}.packed"] = packed.tensor.numpy()
        return tm

    # AWQ 'gemm'
    fmt = wf.AWQFormat(block_in=G)
    raw = {".qweight": rnd_i32(K, N // 8), ".qzeros": rnd_i32(K // G, N // 8), ".scales": rnd_f16(K // G, N)}
    tm = save("awq", fmt, raw, fmt.suffix_map)
    out["awq.dequant"] = fmt.dequant({k: v.to(torch.float32) if k == "weight" else v.to(torch.float32)
                                      for k, v in tm.items()}, None)["weight"].numpy()

    # GPTQ (sym, no act-order): with and without a stored qzeros tensor
    fmt = wf.GPTQFormat(block_in=G)
    raw = {".qweight": rnd_i32(K // 8, N), ".qzeros": rnd_i32(K // G, N // 8), ".scales": rnd_f16(K // G, N)}
    tm = save("gptq", fmt, raw, fmt.suffix_map)
    out["gptq.synth_zeros"] = fmt.synthesize_zeros(tm["scales"]).numpy()

    # compressed-tensors pack-quantized int4
    fmt = wf.CompressedTensorFormat(block_in=G)
    raw = {".weight_packed": rnd_i32(N, K // 8), ".weight_scale": rnd_f16(N, K // G),
           ".weight_zero_point": rnd_i32(N // 8, K // G)}
    tm = save("ct", fmt, raw, fmt.suffix_map)
    out["ct.synth_zeros"] = fmt.synthesize_zeros(tm["scales"]).numpy()
    out["ct.dequant"] = fmt.dequant({"weight": tm["weight"], "scales": tm["scales"].to(torch.float32),
                                     "zeros": tm["zeros"].to(torch.float32)}, None)["weight"].numpy()

    # FP8 block-scaled (weights travel as raw bytes)
    fmt = wf.FP8Format()
    K8, N8 = 256, 384
    raw = {".weight": torch.randint(0, 256, (N8, K8), generator=g, dtype=torch.uint8),
           ".weight_scale_inv": (torch.rand(N8 // 128, K8 // 128, generator=g) * 0.01 + 1e-4)}
    raw[".weight"][raw[".weight"] == 0x7F] = 0x7E          # no NaN payloads
    raw[".weight"][raw[".weight"] == 0xFF] = 0xFE
    for suffix, t in raw.items():
        out[f"fp8.in{suffix}"] = t.numpy()
    for suffix, kind in fmt.suffix_map.items():
        if suffix in raw:
            out[f"fp8.tm.{kind}"] = fmt.normalize(raw[suffix], kind).contiguous().numpy()

    # accepts() truth table on the suffix sets above (format detection)
    cands = [wf.AWQFormat(block_in=G), wf.GPTQFormat(block_in=G), wf.CompressedTensorFormat(block_in=G), wf.FP8Format(),
             wf.TrivialFormat()]
    cases = {
        "awq": {".qweight": rnd_i32(K, N // 8), ".scales": rnd_f16(K // G, N), ".qzeros": rnd_i32(K // G, N // 8)},
        "gptq": {".qweight": rnd_i32(K // 8, N), ".scales": rnd_f16(K // G, N), ".qzeros": rnd_i32(K // G, N // 8)},
        "ct": {".weight_packed": rnd_i32(N, K // 8), ".weight_scale": rnd_f16(N, K // G)},
        "fp8": {".weight": torch.zeros(N, K, dtype=torch.uint8), ".weight_scale_inv": torch.ones(1, 2)},
        "dense": {".weight": torch.zeros(N, K, dtype=torch.float16)},
    }
    table = np.array([[int(bool(c.accepts(av))) for c in cands] for av in cases.values()], np.int8)
    out["accepts.table"] = table
    out["accepts.cases"] = np.array(list(cases.keys()))
    out["accepts.formats"] = np.array([c.name for c in cands])

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out), "arrays")
    print(table)


if __name__ == "__main__":
    main()
