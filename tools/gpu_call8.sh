#!/bin/bash
# Round-2 call 8 (1 GPU): rotated de-quantisation loop + 3 A slots + short-K heuristic: parity, timeline, timing.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 6 "gpurun_out/$name.log" | cut -c1-400; }
run linear_tests 600 python -m pytest tests/test_linear_gpu.py tests/test_linear_fp8_gpu.py tests/test_decode_step_gpu.py tests/test_decode_step_variants_gpu.py -x -q -m gpu
B200_RUN_UNVERIFIED=1 run pending 600 python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -p no:cacheprovider
run gemm_trace_roles 120 build/gemm_trace 64 20 1
run ref_gemm_bench 300 python tools/ref_gemm_bench.py 64 20
run gemm_sweep_7b 300 build/gemm_sweep 7b 64 20
run gemm_sweep_7b-tp8 300 build/gemm_sweep 7b-tp8 64 20
run bench 600 python bench.py
