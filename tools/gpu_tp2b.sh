#!/bin/bash
# NVLS (multimem.ld_reduce) fused all-reduce at N = 2: parity test and bench, next to the peer-load kernel.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 5 "gpurun_out/$name.log" | cut -c1-500; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
B200_TP_NVLS=1 run tp2_test_nvls 600 python -m pytest tests/test_tp_gpu.py -x -q -m gpu
B200_TP_NVLS=1 run tp2_bench_nvls 600 $TR --master-port 29531 bench.py --gpus 2 --steps 64 --warmup 8 --no-cpu-baseline
run tp2_bench_ipc 600 $TR --master-port 29532 bench.py --gpus 2 --steps 64 --warmup 8 --no-cpu-baseline
B200_TP_NVLS=1 B200_STEP_TRACE=1 B200_ENGINE_GRAPHS=0 run tp2_trace_nvls 600 $TR --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 --no-graph --no-cpu-baseline
B200_STEP_TRACE=1 B200_ENGINE_GRAPHS=0 run tp2_trace_ipc 600 $TR --master-port 29534 bench.py --gpus 2 --steps 4 --warmup 3 --no-graph --no-cpu-baseline
