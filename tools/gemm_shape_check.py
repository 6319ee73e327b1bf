"""Development tool: u4 GEMM vs the CPU oracle over a grid of (K, N) with the plan the library picks (debugging aid)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from lmdeploy_b200 import _lib as L
from oracle import linear as OL
from oracle.numerics import F16, rn
from test_linear_gpu import DevLinear, _mk_u4

lib = L.load()
ms = [int(a) for a in sys.argv[1:]] or [16]
for K in (1024, 1536, 2048, 2560, 3072, 3584, 4096):
    for N in (4096, 6144):
        rng = np.random.default_rng(K + N)
        q, s, z = _mk_u4(rng, K, N)
        lin = DevLinear(lib, K, N, 1, F16, q, s, z)
        wd = OL.dequant_u4(q, s, z, F16)
        for m in ms:
            x = rn(rng.standard_normal((m, K)).astype(np.float32), F16)
            plan = (C.c_int * 8)()
            lib.b200_linear_plan(C.byref(lin.d), m, plan)
            got = lin.forward(x)
            ref = OL.linear(x, wd, F16)
            err = np.abs(got - ref)
            bad = err > 2e-4 + 2e-3 * np.abs(ref)
            cols = np.unique(np.nonzero(bad)[1])
            print(f"K={K} N={N} m={m} plan tiles={plan[0]} splits={plan[1]} nw={plan[2]} nx={plan[3]} na={plan[4]}: max err {err.max():.4f} "
                  f"bad {bad.sum()}/{bad.size} cols {cols[:6]}..{cols[-3:] if cols.size else ''} (n bad cols {cols.size})", flush=True)
