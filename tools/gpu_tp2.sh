#!/bin/bash
# Round-2 two-GPU call: TP parity test, NVLS probe, bench at N=2 (raw lines kept).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; tail -n 8 "gpurun_out/$name.log"; }
nvidia-smi topo -m > gpurun_out/tp2_topo.log 2>&1
run tp2_test 600 python -m pytest tests/test_tp_gpu.py -x -q -m gpu
run tp2_nvls_probe 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nvls_probe.py
run tp2_bench 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 64 --warmup 8
B200_STEP_TRACE=1 run tp2_trace 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-graph
