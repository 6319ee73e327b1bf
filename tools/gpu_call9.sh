#!/bin/bash
# Round-2 call 9 (1 GPU): attention producer with L2 prefetch (parity + timing), the reference's own linear smoke suite against the
# _turbomind shim (needs a scratch copy of the reference's tests/turbomind/linear under build/reftests/, removed afterwards).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 8 "gpurun_out/$name.log" | cut -c1-400; }
run attn_tests 600 python -m pytest tests/test_decode_attention_gpu.py tests/test_ref_attention_gpu.py tests/test_decode_step_gpu.py -x -q -m gpu
run attn_variants 300 python tools/attn_variants.py 20
B200_REF_LINEAR_TESTS=$PWD/build/reftests/turbomind_linear run ref_linear_harness 900 python -m pytest tests/test_reference_linear_harness_gpu.py -q -m gpu -s
run linear_tests 600 python -m pytest tests/test_linear_gpu.py -x -q -m gpu
run ref_gemm_bench 300 python tools/ref_gemm_bench.py
run bench 600 python bench.py --no-cpu-baseline
run bench_c3 600 python bench.py --config c3 --no-cpu-baseline
