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
    # per-head q/k RMSNorm: argument checks happen before any device work
    assert lib.b200_qk_rmsnorm(None, 1024, None, None, 4, 4, 2, 128, 1e-6, 0, None) == L.ERR_INVALID_ARG
    assert lib.b200_qk_rmsnorm(8, 1024, 8, 8, 4, 4, 2, 64, 1e-6, 0, None) == L.ERR_UNSUPPORTED and b"head_dim" in lib.b200_last_error()
    assert lib.b200_qk_rmsnorm(8, 512, 8, 8, 4, 4, 2, 128, 1e-6, 0, None) == L.ERR_INVALID_ARG           # stride < (4 + 2) * 128
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


def _prototypes():
    """name -> (return type text, [parameter texts]) of every B200_API declaration in the header."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200_turbomind.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"B200_API\s+([\w\s\*]+?)\b(b200_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        out[name] = (" ".join(ret.split()), params)
    return out


def test_ctypes_signatures_match_the_header_prototypes(lib):
    """Every function the binding declares argtypes for takes as many parameters as the header says, pointer parameters are
    declared as pointers (never a bare int: ctypes would truncate a 64-bit address), and non-int returns carry a restype."""
    import ctypes as C
    protos = _prototypes()
    assert len(protos) >= 30
    declared = 0
    for name, (ret, params) in protos.items():
        fn = getattr(lib, name)
        if ret in ("size_t", "const char*", "const char *"):
            assert fn.restype is not C.c_int, f"{name} returns {ret} but the binding leaves restype at c_int"
        if fn.argtypes is None:
            # undeclared argtypes pass Python ints as 32-bit: only acceptable when nothing pointer-like is passed
            assert not any("*" in p or "[" in p for p in params), f"{name} takes pointers but the binding declares no argtypes"
            continue
        declared += 1
        assert len(fn.argtypes) == len(params), (name, params, fn.argtypes)
        for p, t in zip(params, fn.argtypes):
            is_ptr = "*" in p or "[" in p            # an array parameter decays to a pointer
            t_ptr = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) == "P"
            assert is_ptr == bool(t_ptr), (name, p, t)
            if not is_ptr:
                want = {"int": C.c_int, "float": C.c_float, "size_t": C.c_size_t, "int64_t": C.c_int64}[p.split()[-2] if len(p.split()) > 1 else p]
                assert t is want, (name, p, t)
    assert declared >= 25
