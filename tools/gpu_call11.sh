#!/bin/bash
# Round-2 call 11 (1 GPU): u4 GEMM with the de-quantisation loop unrolled over one lap of the W ring and the f16 "subnormal"
# weight layout: parity (linear tests, reference GEMM, whole step vs reference kernels), then timing with the layout on / off.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 8 "gpurun_out/$name.log" | cut -c1-330; }
run linear_tests 900 python -m pytest tests/test_linear_gpu.py tests/test_linear_fp8_gpu.py -x -q -m gpu
B200_RUN_UNVERIFIED=1 run pending 900 python -m pytest tests/test_zz_pending_gpu.py -x -q -m gpu -s
run step_tests 900 python -m pytest tests/test_decode_step_gpu.py tests/test_decode_step_variants_gpu.py tests/test_engine_gpu.py -x -q -m gpu
run ref_gemm_bench 300 python tools/ref_gemm_bench.py
B200_U4_PLAIN=1 run ref_gemm_bench_plain 300 python tools/ref_gemm_bench.py
run bench 600 python bench.py --no-cpu-baseline
B200_U4_PLAIN=1 run bench_plain 600 python bench.py --no-cpu-baseline
