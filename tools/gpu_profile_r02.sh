#!/bin/bash
# Round-2 ncu evidence (1 GPU): launch list of one un-captured decode step, `--set full` of the dominant kernel (decode attention)
# and of the four u4 GEMMs.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
K='gemm_kernel|decode_attention_kernel|decode_reduce|kv_prologue|rmsnorm|argmax|embedding|rope_table|plan_kernel|step_advance'
B="python bench.py --no-graph --no-cpu-baseline --steps 2 --warmup 3 --attn-iters 1 --engine-ingest-limit 0"
# 32 prefill prologues + 3 warm-up steps x 296 launches are skipped; one whole step (296 launches) is listed
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 920 -c 296 --csv \
    --log-file gpurun_out/launches_r02.csv $B > gpurun_out/profile_launches.log 2>&1
echo "launch list exit $?"; tail -n 2 gpurun_out/profile_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_attention_kernel -s 100 -c 1 \
    -o gpurun_out/r02_attn -f $B > gpurun_out/profile_attn.log 2>&1
echo "attention capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 400 -c 4 \
    -o gpurun_out/r02_gemm -f $B > gpurun_out/profile_gemm.log 2>&1
echo "gemm capture exit $?"
ls -la gpurun_out/*.ncu-rep
