#!/bin/bash
# Round-2 final eight-GPU call: all-reduce microbenchmark (peer loads vs NVLS) at N = 8 and 4, C2 at N = 8 (both exchanges), 4, 2,
# C5 (Llama-3-70B FP8, TP 8).  Raw lines kept under gpurun_out/ -> profiles/r02_raw/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" "gpurun_out/$name.log" | tail -n 3 | cut -c1-700; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run f_comm8_ipc 200 $TR --nproc-per-node 8 --master-port 29551 tools/comm_bench.py --short
B200_TP_NVLS=1 run f_comm8_nvls 200 $TR --nproc-per-node 8 --master-port 29552 tools/comm_bench.py --short
run f_comm4_ipc 200 $TR --nproc-per-node 4 --master-port 29553 tools/comm_bench.py --short
B200_TP_NVLS=1 run f_comm4_nvls 200 $TR --nproc-per-node 4 --master-port 29554 tools/comm_bench.py --short
run f_tp8_c2_ipc 400 $TR --nproc-per-node 8 --master-port 29555 bench.py --gpus 8 --steps 64 --warmup 8
B200_TP_NVLS=1 run f_tp8_c2_nvls 400 $TR --nproc-per-node 8 --master-port 29556 bench.py --gpus 8 --steps 64 --warmup 8
run f_tp4_c2_ipc 400 $TR --nproc-per-node 4 --master-port 29557 bench.py --gpus 4 --steps 64 --warmup 8
run f_tp2_c2_ipc 400 $TR --nproc-per-node 2 --master-port 29558 bench.py --gpus 2 --steps 64 --warmup 8
B200_TP_NVLS=1 run f_tp8_c5_nvls 600 $TR --nproc-per-node 8 --master-port 29559 bench.py --gpus 8 --config c5 --steps 32 --warmup 4
