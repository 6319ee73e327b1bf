#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 8 "gpurun_out/$name.log"; }
run attn_variants 300 python tools/attn_variants.py 20
run attn_variants_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention|prologue|reduce|plan_kernel" --csv --log-file gpurun_out/attn_variants_launches.csv python tools/attn_variants.py 4
export B200_RUN_UNVERIFIED=1
run pending 600 python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -p no:cacheprovider -s
