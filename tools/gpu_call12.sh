#!/bin/bash
# Round-2 call 12 (1 GPU): attention main kernel with early KV streaming (producer warp ahead of griddepcontrol.wait):
# parity with it on, then the decode step with it on / off.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 6 "gpurun_out/$name.log" | cut -c1-330; }
B200_ATTN_EARLY=1 run early_attn_tests 600 python -m pytest tests/test_decode_attention_gpu.py tests/test_ref_attention_gpu.py -x -q -m gpu
B200_ATTN_EARLY=1 run early_step_tests 900 python -m pytest tests/test_decode_step_gpu.py tests/test_decode_step_variants_gpu.py tests/test_engine_gpu.py -x -q -m gpu
B200_ATTN_EARLY=1 B200_RUN_UNVERIFIED=1 run early_l2 600 python -m pytest tests/test_zz_pending_gpu.py -x -q -m gpu -k greedy_tokens -s
for rep in 1 2; do
B200_ATTN_EARLY=1 run bench_early_$rep 600 python bench.py --no-cpu-baseline
B200_ATTN_EARLY=0 run bench_noearly_$rep 600 python bench.py --no-cpu-baseline
done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_early_*.log gpurun_out/bench_noearly_*.log
