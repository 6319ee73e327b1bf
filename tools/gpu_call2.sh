#!/bin/bash
# Round-2 second GPU call: diagnose the reference-GEMM illegal instruction, try the PTX-JIT build, run the ungated suite,
# dump the GEMM per-role timeline, run the bench with the engine e2e arm.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 6 "gpurun_out/$name.log"; }
run suite 900 python -m pytest tests -x -q -m gpu
export B200_RUN_UNVERIFIED=1
P="python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -p no:cacheprovider"
B200_REF_GEMM_LIB=ptx run ref_gemm_ptx 600 $P -x -k "reference_gemm and 1-shape0"
run ref_gemm_sanitizer 600 compute-sanitizer --tool memcheck --print-limit 5 $P -x -k "reference_gemm and 1-shape0"
B200_REF_GEMM_LIB=ptx run ref_gemm_ptx_all 600 $P -k "reference_gemm"
B200_REF_GEMM_LIB=ptx run l2_ptx 600 $P -s -k "assembled_from_reference_kernels"
unset B200_RUN_UNVERIFIED
run gemm_trace_roles 120 build/gemm_trace 64 20 1
run bench 600 python bench.py
B200_REF_GEMM_LIB=ptx B200_BENCH_REF_GEMM=1 run bench_ref_gemm 600 python bench.py --no-cpu-baseline --steps 16 --warmup 4
