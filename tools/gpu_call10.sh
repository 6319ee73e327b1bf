#!/bin/bash
# Round-2 call 10 (1 GPU): whole-step time with the three attention builds (HEAD / post-scaled exact-int / post-scaled subnormal),
# prefetch off; whole-step parity vs the reference kernels with the default build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/lmdeploy_b200/lib
for rep in 1 2; do
for lib in libb200tm_head.so libb200tm_exactint.so libb200tm.so; do
  echo "=== bench $lib rep $rep"
  B200_LIB_PATH=$L/$lib timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'attn main us', d['roofline']['us_per_launch'], 'full op us', d['roofline']['full_op']['us_per_call'], 'ref', d.get('reference_kernel'))
"
done
done 2>&1 | tee gpurun_out/bench_ab.log
echo "=== L2 parity"
B200_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -k "greedy_tokens" -s 2>&1 | tail -n 5 | tee gpurun_out/l2_postscale.log
