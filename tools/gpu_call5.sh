#!/bin/bash
# Round-2 call 5: attention variants (3-kernel / prologue fused / fully fused) parity + step time; GEMM after the epilogue revert.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 6 "gpurun_out/$name.log"; }
run attn_tests 600 python -m pytest tests/test_decode_attention_gpu.py tests/test_ref_attention_gpu.py -x -q -m gpu
run suite 900 python -m pytest tests -x -q -m gpu
export B200_RUN_UNVERIFIED=1
run pending 600 python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -p no:cacheprovider -s
unset B200_RUN_UNVERIFIED
run ref_gemm_bench 300 python tools/ref_gemm_bench.py 64 20
run bench 600 python bench.py
B200_ATTN_FUSED=0 run bench_attn0 600 python bench.py --no-cpu-baseline --steps 32 --warmup 4
B200_ATTN_FUSED=3 run bench_attn3 600 python bench.py --no-cpu-baseline --steps 32 --warmup 4
