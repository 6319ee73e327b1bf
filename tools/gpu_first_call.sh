#!/bin/bash
# First GPU call of a round: everything that was written or changed without a device gets its first run, each step under
# its own timeout so one hang cannot eat the call.  Outputs land in gpurun_out/.
#   bash tools/build_tools.sh && gpurun --timeout 2700 -- 'bash tools/gpu_first_call.sh'   (typically ~15 min; every step has its own limit)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { # name, timeout seconds, command...
    local name=$1 t=$2; shift 2
    echo "=== $name"
    timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
    echo "exit $? ($name)"; tail -n 6 "gpurun_out/$name.log"
}
run verified_suite 900 python -m pytest tests -x -q -m gpu
# pending tests in separate processes, least risky first, so one hanging kernel costs one group and not the rest
export B200_RUN_UNVERIFIED=1
P="python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -p no:cacheprovider"
run pending_fp8_linear 240 $P -k "fp8_per_column or fp8_block_scales or fp8_gated"
run pending_ref_gemm 240 $P -k "reference_gemm"
run pending_ref_norm 120 $P -k "rmsnorm_matches_reference"
run pending_l2_reference_step 300 $P -s -k "assembled_from_reference_kernels"
run pending_bf16_steps 300 $P -k "bf16_matches_oracle or synthetic_fp8 or qkv_bias or odd_gqa"
run pending_qk_norm 240 $P -k "qk_rmsnorm or qk_norm"
run pending_prefill_oracle 420 $P -k "prefill_attention_matches_oracle or prefill_single or prefill_attention_bf16"
run pending_prefill_reference 240 $P -k "prefill_attention_matches_reference"
run pending_chunk_step 300 $P -k "chunk_step"
run pending_engine 420 $P -k "engine"
unset B200_RUN_UNVERIFIED
[ -x build/gemm_sweep ] && run gemm_sweep_7b 240 build/gemm_sweep 7b 64 20
[ -x build/gemm_sweep ] && run gemm_sweep_70b 120 build/gemm_sweep 70b-tp8 32 20
[ -x build/gemm_sweep ] && run gemm_sweep_7b_4tiles 120 build/gemm_sweep 7b 64 20 4     # tool-only 4-tile kernel, own process
[ -x build/gemm_trace ] && run gemm_trace_wall 120 build/gemm_trace 64 20 0     # wall-clock anatomy of back-to-back GEMM launches
[ -x build/tma_issue_rate ] && run tma_issue_rate 120 build/tma_issue_rate
[ -x build/umma_probe ] && run umma_probe 60 build/umma_probe
run ref_gemm_bench 300 python tools/ref_gemm_bench.py 64 20
run bench 420 python bench.py
B200_BENCH_REF_GEMM=1 run bench_with_ref_gemm 420 python bench.py --no-cpu-baseline --steps 16 --warmup 4
run bench_c3 420 python bench.py --config c3 --no-cpu-baseline --steps 16 --warmup 4    # Llama-3-8B fp16 + KV-int4, bs 128, 4K ctx
B200_BENCH_SEQUENTIAL_PAGES=1 run bench_sequential_pages 300 python bench.py --no-cpu-baseline --steps 16 --warmup 4
# the sweep's winners as a tuned plan table, and the step with it
cat gpurun_out/gemm_sweep_7b.log gpurun_out/gemm_sweep_70b.log 2>/dev/null | grep '^PLAN' > gpurun_out/plans.txt
if [ -s gpurun_out/plans.txt ]; then
    B200_GEMM_PLANS=gpurun_out/plans.txt run bench_tuned_plans 420 python bench.py --no-cpu-baseline
fi
