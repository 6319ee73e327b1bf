#!/bin/bash
# Round-2 eight-GPU call: NVLS probe, TP test, C2 at N=8 / 4, C5 (Llama-3-70B FP8, TP8), C4 (InternLM2-20B, TP4 / TP2); raw lines kept.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 4 "gpurun_out/$name.log" | cut -c1-600; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run tp8_nvls_probe 200 $TR --nproc-per-node 8 --master-port 29521 tools/nvls_probe.py
run tp8_bench_c2 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 64 --warmup 8
run tp4_bench_c2 400 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 64 --warmup 8
run tp8_bench_c5 600 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --config c5 --steps 32 --warmup 4
run tp4_bench_c4 500 $TR --nproc-per-node 4 --master-port 29525 bench.py --gpus 4 --config c4 --steps 32 --warmup 4
