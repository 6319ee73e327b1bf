"""The ctypes mirror (lmdeploy_b200/_lib.py) against the C header, by compiling the header: sizeof and every field's
offsetof must agree, field names must exist in the C struct, and the enums must carry the same values.  A binding that
drifts from include/b200_turbomind.h corrupts the step silently on the GPU; this catches it on the CPU."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lmdeploy_b200 import _lib as L  # noqa: E402

PAIRS = [("b200_rope_config", L.RopeConfig), ("b200_attention_params", L.AttentionParams),
         ("b200_linear_desc", L.LinearDesc), ("b200_model_config", L.ModelConfig), ("b200_step_inputs", L.StepInputs),
         ("b200_sampling_params", L.SamplingParams)]
ENUMS = [("B200_T_EMBEDDING", L.T_EMBEDDING), ("B200_T_FINAL_NORM", L.T_FINAL_NORM), ("B200_T_ATTN_NORM", L.T_ATTN_NORM),
         ("B200_T_FFN_NORM", L.T_FFN_NORM), ("B200_T_QKV_BIAS", L.T_QKV_BIAS), ("B200_ERR_INVALID_ARG", L.ERR_INVALID_ARG),
         ("B200_ERR_UNSUPPORTED", L.ERR_UNSUPPORTED), ("B200_ERR_WORKSPACE", L.ERR_WORKSPACE), ("B200_ERR_STATE", L.ERR_STATE)]


def test_ctypes_structures_match_the_header(tmp_path):
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "b200_turbomind.h"', 'int main(void) {']
    for cname, cls in PAIRS:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, ftype in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname}*)0)->{fname}));')
    for ename, _ in ENUMS:
        lines.append(f'  printf("{ename} %d\\n", (int){ename});')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr          # a field the header does not have fails to compile, by name
    got = dict((ln.split()[0], [int(x) for x in ln.split()[1:]]) for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in PAIRS:
        assert got[cname] == [C.sizeof(cls)], (cname, got[cname], C.sizeof(cls))
        covered = 0
        for fname, ftype in cls._fields_:
            f = getattr(cls, fname)
            assert got[f"{cname}.{fname}"] == [f.offset, f.size], (cname, fname, got[f"{cname}.{fname}"], f.offset, f.size)
            covered = max(covered, f.offset + f.size)
        # no C field beyond the last mirrored one (only tail padding may follow)
        assert C.sizeof(cls) - covered < C.alignment(cls), (cname, "the header has fields the mirror lacks")
    for ename, val in ENUMS:
        assert got[ename] == [val], ename


def test_the_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 and as C++ with -Wall -Werror, on its own."""
    for compiler, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        r = subprocess.run([compiler, std, "-Wall", "-Werror", "-fsyntax-only", "-x", "c" if compiler == "gcc" else "c++",
                            os.path.join(ROOT, "include", "b200_turbomind.h")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
