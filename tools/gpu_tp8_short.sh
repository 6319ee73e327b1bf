#!/bin/bash
# Round-2 eight-GPU call (short): C2 at N = 8 with the default exchange (NVLS from 4 ranks up) and with peer loads, then the
# all-reduce microbenchmark both ways.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" "gpurun_out/$name.log" | tail -n 3 | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
run f_tp8_c2_default 150 $TR --master-port 29561 bench.py --gpus 8 --steps 64 --warmup 8
B200_TP_NVLS=0 run f_tp8_c2_ipc 150 $TR --master-port 29562 bench.py --gpus 8 --steps 64 --warmup 8
B200_TP_NVLS=1 run f_comm8_nvls 100 $TR --master-port 29563 tools/comm_bench.py --short
B200_TP_NVLS=0 run f_comm8_ipc 100 $TR --master-port 29564 tools/comm_bench.py --short
