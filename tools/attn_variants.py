"""Development tool (measurement only): the decode attention op at the C2 shape in its three forms -- debug_flags 4 (prologue
kernel + main kernel + fix-up kernel), 0 (the process default, B200_ATTN_FUSED), 8 (everything in one kernel) -- CUDA-event
timed back to back, with and without a per-step rope table.  Run it under `ncu --metrics gpu__time_duration.sum` for the
per-kernel split."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lmdeploy_b200 import _lib as L
from lmdeploy_b200.turbomind import SPECS, DecodeModel
from lmdeploy_b200.engine import StaticBatchRunner

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
spec = SPECS["internlm2-7b"]
import dataclasses
spec = dataclasses.replace(spec, layers=4)
m = DecodeModel(spec, quant_policy=8, max_batch=64)
m.load_synthetic(0)
r = StaticBatchRunner(m, batch=64, max_len=2048 + 16, seed=0)
r.prefill_random(2048)
r.time_attention_kernel(2049, iters=2)
p = r._attn_params[0]
cu_k = r._attn_params[4]
table = torch.zeros(64 * 64 * 2, dtype=torch.float32, device="cuda")
L.check(m.lib.b200_rope_table(table.data_ptr(), C.byref(m.cfg.rope), r.cu_q.data_ptr(), cu_k.data_ptr(), 64, 64,
                              torch.cuda.current_stream().cuda_stream), "rope_table")
plan = torch.zeros(65, dtype=torch.int32, device="cuda")
L.check(m.lib.b200_decode_attention_plan(plan.data_ptr(), cu_k.data_ptr(), None, 64, torch.cuda.current_stream().cuda_stream), "plan")
for with_table in (0, 1):
    p.rope_table = table.data_ptr() if with_table else None
    p.plan = plan.data_ptr() if with_table else None
    for flags in (4, 0, 8):
        p.debug_flags = flags
        for i in range(3):
            p.layer_offset = (i % 4) * m.layer_bytes
            L.check(m.lib.b200_decode_attention(C.byref(p)), "attn")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            p.layer_offset = (i % 4) * m.layer_bytes
            L.check(m.lib.b200_decode_attention(C.byref(p)), "attn")
        e1.record()
        torch.cuda.synchronize()
        print(f"table+plan={with_table} flags={flags}: {e0.elapsed_time(e1) * 1e3 / iters:8.2f} us per call", flush=True)
