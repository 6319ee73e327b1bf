import sys, ctypes as C, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from oracle import attention as OA, ref_lib
from harness import PagedCase, from_dev_T
from lmdeploy_b200 import _lib as L
lib = L.load(); ref = ref_lib.load()
import test_ref_attention_gpu as T
for kv_bits, (hq, hkv) in [(8, (32, 8)), (4, (32, 8)), (8, (8, 8))]:
    hist = [0, 1, 63, 64, 65, 200, 777, 1500]
    case = PagedCase(batch_hist=hist, num_heads=hq, num_kv_heads=hkv, kv_bits=kv_bits, seed=kv_bits + hq, exact_math=False)
    oo, po, orf, pr = T._run_both(case, lib, ref)
    oracle = case.run_oracle()
    print("kv", kv_bits, "heads", hq, hkv)
    e_or = np.abs(oo - oracle).reshape(len(hist), -1).max(1); e_rr = np.abs(orf - oracle).reshape(len(hist), -1).max(1)
    e = np.abs(oo - orf).reshape(len(hist), -1).max(1)
    print(" per-seq max |ours-oracle|", np.round(e_or, 5)); print(" per-seq max |ref-oracle| ", np.round(e_rr, 5)); print(" per-seq max |ours-ref|   ", np.round(e, 5))
    # byte diffs by region for the appended token of each sequence
    lay = case.layout; D = 128; rb = D * kv_bits // 8
    for b, h in enumerate(hist):
        blk = case.block_tables[b][h // 64]; bt = h % 64
        base = case.cache.layer_offset(case.layer)
        for name, pool in (("ours", po), ("ref", pr)):
            pass
        nk = nv = npar = 0
        for hh in range(hkv):
            kd = base + (2 * hh) * 64 * rb + bt * rb; vd = kd + 64 * rb
            kp = base + hkv * 2 * 64 * rb + (2 * hh) * 256 + bt * 4; vp = kp + 256
            nk += (po[blk, kd:kd + rb] != pr[blk, kd:kd + rb]).sum(); nv += (po[blk, vd:vd + rb] != pr[blk, vd:vd + rb]).sum()
            npar += (po[blk, kp:kp + 4] != pr[blk, kp:kp + 4]).sum() + (po[blk, vp:vp + 4] != pr[blk, vp:vp + 4]).sum()
        print(f"  seq {b} hist {h}: differing bytes K {nk} V {nv} params {npar}")
    print(" total differing bytes", (po != pr).sum(), "ours vs oracle cache", (po != case.cache.pool).sum(), "ref vs oracle cache", (pr != case.cache.pool).sum())
