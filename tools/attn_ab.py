"""Development tool (measurement only): CUDA-event time of the decode attention main kernel and of the whole op at the C2
shape (bs 64, ctx 2048, KV int8, f16) for the library named by B200_LIB_PATH and the prefetch distance in
B200_ATTN_PREFETCH.  tools/attn_ab.sh runs it once per variant in a fresh process."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import torch
from lmdeploy_b200.turbomind import SPECS, DecodeModel
from lmdeploy_b200.engine import StaticBatchRunner

kv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
spec = dataclasses.replace(SPECS["internlm2-7b"], layers=8)
m = DecodeModel(spec, quant_policy=kv, max_batch=64)
m.load_synthetic(0)
r = StaticBatchRunner(m, batch=64, max_len=2048 + 16, seed=0)
r.prefill_random(2048)
res = []
for rep in range(3):
    a = r.time_attention_kernel(2049, iters=iters)
    b = r.time_attention_kernel(2049, iters=iters, full=True)
    res.append((a["us"], b["us"]))
tag = f"lib={os.path.basename(os.environ.get('B200_LIB_PATH', 'libb200tm.so'))} prefetch={os.environ.get('B200_ATTN_PREFETCH', '0')} kv={kv}"
print(tag, " main-only us:", " ".join(f"{x[0]:.1f}" for x in res), "  full op us:", " ".join(f"{x[1]:.1f}" for x in res), flush=True)
