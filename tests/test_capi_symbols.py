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


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (it is test infrastructure)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lmdeploy_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
