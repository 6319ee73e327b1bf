"""Host logic of the GEMM launcher (no GPU): the launch plan b200_linear_plan returns must respect the hardware
budgets the kernel relies on, for every shape of the hot path and for a random sweep."""
import ctypes as C

import numpy as np
import pytest

from lmdeploy_b200 import _lib as L

SMS, SMEM_MAX, TMEM_COLS = 148, 227 * 1024 - 2048, 512
W_BLOCK = {0: 32768, 1: 8704, 2: 16400, 3: 16896}      # bytes of one (n-tile, k-group) of prepared weights per format


def _plan(lib, K, N, fmt, m):
    d = L.LinearDesc(K, N, fmt, 0, 128, None, None)
    out = (C.c_int * 8)()
    rc = lib.b200_linear_plan(C.byref(d), m, out)
    assert rc == 0, lib.b200_last_error()
    return dict(zip(("tiles", "splits", "nw", "nx", "na", "tmem_cols", "a_col", "smem"), list(out)))


def _check(p, K, N, fmt, m):
    mt = max(16, (m + 15) // 16 * 16)
    kgs, rows = K // 128, 128 * p["tiles"]
    assert p["tiles"] in (1, 2) and N % rows == 0
    assert fmt != 0 or p["tiles"] == 1                          # dense: one tile (static_assert in the kernel)
    ctas = N // rows * p["splits"]
    assert p["splits"] in (1, 2, 4, 8)
    if p["splits"] > 1:
        assert ctas <= SMS, "split-K CTAs wait for each other: they must be co-resident (one CTA per SM)"
        assert kgs // p["splits"] >= 4
    wslot = ((W_BLOCK[fmt] + 1023) & ~1023) * p["tiles"]
    assert p["smem"] == p["nw"] * wslot + p["nx"] * mt * 256 <= SMEM_MAX
    assert 3 <= p["nw"] <= 8 and 2 <= p["nx"] <= 6
    d_cols = (64 if mt <= 64 else 128) * p["tiles"]
    if fmt == 0:
        assert p["tmem_cols"] >= mt and p["tmem_cols"] in (32, 64, 128)
    else:
        assert p["a_col"] >= d_cols and p["a_col"] + 64 * p["na"] <= p["tmem_cols"] <= TMEM_COLS
        assert p["na"] % p["tiles"] == 0 and p["na"] // p["tiles"] >= 1


HOT_PATH = [  # (K, N, fmt) of the BASELINE.json configs, TP 1..8
    (4096, 6144, 1), (4096, 4096, 1), (4096, 28672, 1), (14336, 4096, 1), (4096, 92544, 0),      # InternLM2-7B
    (2048, 3072, 1), (2048, 4096, 1), (4096, 14336, 1), (7168, 4096, 1),                          # ... TP2
    (4096, 768, 1), (512, 4096, 1), (4096, 3584, 1), (1792, 4096, 1), (4096, 11648, 0),          # ... TP8
    (6144, 8192, 1), (6144, 6144, 1), (6144, 32768, 1), (16384, 6144, 1),                         # InternLM2-20B
    (4096, 6144, 0), (4096, 4096, 0), (4096, 28672, 0), (14336, 4096, 0), (4096, 128256, 0),     # Llama-3-8B fp16
    (8192, 1280, 2), (1024, 8192, 2), (8192, 7168, 2), (3584, 8192, 2),                           # Llama-3-70B fp8, TP8
]


@pytest.mark.parametrize("m", [1, 16, 64, 128])
def test_plan_budgets_hot_path(lib, m):
    for K, N, fmt in HOT_PATH:
        _check(_plan(lib, K, N, fmt, m), K, N, fmt, m)


def test_plan_budgets_random_sweep(lib):
    rng = np.random.default_rng(0)
    for _ in range(400):
        K, N = 128 * int(rng.integers(1, 160)), 128 * int(rng.integers(1, 800))
        fmt, m = int(rng.integers(0, 3)), int(rng.integers(1, 129))
        _check(_plan(lib, K, N, fmt, m), K, N, fmt, m)


def test_plan_is_a_pure_function(lib):
    a = [_plan(lib, 14336, 4096, 1, 64) for _ in range(3)]
    assert a[0] == a[1] == a[2] and a[0]["tiles"] == 2 and a[0]["splits"] == 8
    assert _plan(lib, 4096, 28672, 1, 64)["splits"] == 1          # 112 two-tile CTAs already fill the wave


# ---------------------------------------------------------------------------------------------------------------
# tuned plan table (counterpart of the reference's GEMM dispatch cache)
# ---------------------------------------------------------------------------------------------------------------
def test_plan_table_overrides_the_heuristic_and_is_validated(lib, tmp_path):
    K, N, fmt, m = 4096, 6144, 1, 64
    base = _plan(lib, K, N, fmt, m)
    assert (base["tiles"], base["splits"]) == (2, 4)
    f = tmp_path / "plans.txt"
    f.write_text("# tuned on a B200\nPLAN 1 64 6144 4096 1 2 0 0\nPLAN 1 64 4096 4096 2 4 5 4\n")
    try:
        assert lib.b200_linear_load_plans(str(f).encode()) == 2
        p = _plan(lib, K, N, fmt, m)
        assert (p["tiles"], p["splits"]) == (1, 2)
        _check_budgets_only(p, K, N, fmt, m)
        p2 = _plan(lib, 4096, 4096, 1, 64)
        assert (p2["tiles"], p2["splits"], p2["nw"], p2["nx"]) == (2, 4, 5, 4)
        # other shapes and token counts are untouched
        assert _plan(lib, K, N, fmt, 16) == _plan_fresh(lib, K, N, fmt, 16, loaded=str(f))
        assert _plan(lib, 14336, 4096, 1, 64)["splits"] == 8
        # an invalid entry rejects the whole file and keeps the current table
        bad = tmp_path / "bad.txt"
        for text in ("PLAN 1 64 6144 4096 2 8 0 0\n",          # 24 groups x 8 splits > 148 co-resident CTAs
                     "PLAN 0 64 4096 4096 2 1 0 0\n",          # dense weights: one tile per CTA
                     "PLAN 1 64 6144 4096 2 4 8 6\n",          # rings beyond shared memory
                     "PLAN 1 60 6144 4096 2 4 0 0\n",          # padded tokens are multiples of 16
                     "PLAN 1 64 6144 4096 2 64 0 0\n",         # more splits than K-groups
                     "garbage\n"):
            bad.write_text(text)
            assert lib.b200_linear_load_plans(str(bad).encode()) == L.ERR_INVALID_ARG, text
            assert _plan(lib, K, N, fmt, m)["splits"] == 2
        assert lib.b200_linear_load_plans(str(tmp_path / "missing.txt").encode()) == L.ERR_INVALID_ARG
    finally:
        assert lib.b200_linear_load_plans(None) == 0
    assert _plan(lib, K, N, fmt, m) == base


def _plan_fresh(lib, K, N, fmt, m, loaded):
    """The heuristic's own plan for a shape the loaded table does not mention."""
    lib.b200_linear_load_plans(None)
    p = _plan(lib, K, N, fmt, m)
    lib.b200_linear_load_plans(loaded.encode())
    return p


def _check_budgets_only(p, K, N, fmt, m):
    """_check without the heuristic-specific expectations (splits power of two, >= 4 K-groups per split)."""
    mt = max(16, (m + 15) // 16 * 16)
    rows = 128 * p["tiles"]
    assert N % rows == 0 and (p["splits"] == 1 or N // rows * p["splits"] <= SMS) and p["splits"] <= K // 128
    wslot = ((W_BLOCK[fmt] + 1023) & ~1023) * p["tiles"]
    assert p["smem"] == p["nw"] * wslot + p["nx"] * mt * 256 <= SMEM_MAX
    d_cols = (64 if mt <= 64 else 128) * p["tiles"]
    assert p["a_col"] >= d_cols and p["a_col"] + 64 * p["na"] <= p["tmem_cols"] <= TMEM_COLS
    assert p["na"] % p["tiles"] == 0 and p["na"] // p["tiles"] >= 1


def test_the_shipped_plan_table_loads_and_every_entry_is_a_plan_the_kernel_accepts(lib):
    """lmdeploy_b200/gemm_plans.txt (measured on a B200 by tools/gemm_sweep.cu) is what DecodeModel loads by default: it must
    parse, pass the loader's budget checks, and change the plan of the shapes it names."""
    import ctypes as C
    import os
    from lmdeploy_b200 import _lib as L
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lmdeploy_b200", "gemm_plans.txt")
    entries = [ln.split()[1:9] for ln in open(path) if ln.startswith("PLAN")]
    assert len(entries) >= 10
    try:
        assert lib.b200_linear_load_plans(path.encode()) == len({tuple(e[:4]) for e in entries}), lib.b200_last_error()
        for fmt, mt, N, K, tiles, splits, nw, nx in (map(int, e) for e in entries):
            d = L.LinearDesc(K, N, fmt, 0 if fmt != 3 else 1, 128, None, None)
            out = (C.c_int * 8)()
            assert lib.b200_linear_plan(C.byref(d), mt, out) == 0
            assert (out[0], out[1], out[2], out[3]) == (tiles, splits, nw, nx), (N, K, list(out))
            assert out[7] <= 225 * 1024
    finally:
        lib.b200_linear_load_plans(None)
