#!/bin/bash
# Round-2 call 7 (1 GPU): sampling + engine-sampling tests, whole suite, GEMM plan sweep incl. per-rank TP shapes.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 6 "gpurun_out/$name.log" | cut -c1-400; }
run sampling_tests 600 python -m pytest tests/test_sampling_gpu.py tests/test_engine_gpu.py -x -q -m gpu
B200_RUN_UNVERIFIED=1 run suite 1200 python -m pytest tests -q -m gpu
for m in 7b 7b-tp2 7b-tp4 7b-tp8 20b-tp2 20b-tp4 70b-tp8; do
  run gemm_sweep_$m 300 build/gemm_sweep $m 64 20
done
run gemm_sweep_70b_m32 300 build/gemm_sweep 70b-tp8 32 20
cat gpurun_out/gemm_sweep_*.log | grep '^PLAN' | sort -u > gpurun_out/plans_r02.txt
wc -l gpurun_out/plans_r02.txt
B200_GEMM_PLANS=gpurun_out/plans_r02.txt run bench_plans 600 python bench.py --no-cpu-baseline
B200_GEMM_PLANS=none run bench_noplans 600 python bench.py --no-cpu-baseline
