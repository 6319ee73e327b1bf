#!/bin/bash
# 2 GPUs: all-reduce + residual + RMSNorm microbenchmark, CUDA-IPC peer loads and NVLS.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
{
timeout 300 $TR --master-port 29541 tools/comm_bench.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -n 12
B200_TP_NVLS=1 timeout 300 $TR --master-port 29542 tools/comm_bench.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -n 12
} | tee gpurun_out/comm_bench_${NG:-2}.log
