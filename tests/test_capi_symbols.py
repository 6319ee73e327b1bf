"""The C-ABI library loads and exports every symbol include/b200_turbomind.h declares (no GPU needed)."""
import os
import re


def test_library_exports_all_declared_symbols(lib):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200_turbomind.h")).read()
    names = re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", hdr)
    assert len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert b"sm_100a" in lib.b200_version()


def test_invalid_arguments_fail_loudly(lib):
    import ctypes as C
    from lmdeploy_b200 import _lib as L
    p = L.AttentionParams()
    p.size_per_head = 64
    rc = lib.b200_decode_attention(C.byref(p))
    assert rc == -2 and b"size_per_head" in lib.b200_last_error()
    p.size_per_head = 128
    p.window_size = 4096
    assert lib.b200_decode_attention(C.byref(p)) == -2
    assert lib.b200_decode_attention(None) == -1
    # the prefill entry validates the same block
    assert lib.b200_prefill_attention(None) == L.ERR_INVALID_ARG
    assert lib.b200_prefill_attention(C.byref(p)) == L.ERR_UNSUPPORTED            # window_size still set
    p.window_size = 0
    p.num_heads, p.num_kv_heads, p.block_len, p.token_num, p.batch_size = 8, 2, 64, 0, 0
    assert lib.b200_prefill_attention(C.byref(p)) == L.ERR_INVALID_ARG and b"token_num" in lib.b200_last_error()
    p.token_num, p.batch_size = 16, 1                                             # no workspace
    assert lib.b200_prefill_attention(C.byref(p)) == L.ERR_WORKSPACE
    assert lib.b200_prefill_attention_workspace_bytes(16, 8) >= 16 * 8 * 128 * 2
    # step entry points refuse NULL handles before touching a device
    assert lib.b200_model_decode_step(None, None, None) == L.ERR_INVALID_ARG
    assert lib.b200_model_chunk_step(None, None, 4, None) == L.ERR_INVALID_ARG
    d = L.LinearDesc(4096, 4096, 7, 0, 128, None, None)                            # unknown weight format
    assert lib.b200_linear_prepared_bytes(C.byref(d)) == 0 and b"weight_format" in lib.b200_last_error()
    d = L.LinearDesc(4096, 4096, 3, 0, 64, None, None)                             # fp8 needs 128-wide K groups
    assert lib.b200_linear_prepared_bytes(C.byref(d)) == 0
    d = L.LinearDesc(4096, 4096, 3, 0, 128, None, None)
    assert lib.b200_linear_prepared_bytes(C.byref(d)) == (4096 // 128) ** 2 * (128 * 128 + 128 * 4)


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (it is test infrastructure)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lmdeploy_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
