#!/bin/bash
# Round-2 four-GPU call: all-reduce microbenchmark (peer loads vs NVLS) and C2 at N = 4 with both exchanges.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $? ($name)"; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" "gpurun_out/$name.log" | tail -n 3 | cut -c1-700; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node ${NG:-4}"
run f_comm${NG:-4}_ipc 200 $TR --master-port 29553 tools/comm_bench.py --short
B200_TP_NVLS=1 run f_comm${NG:-4}_nvls 200 $TR --master-port 29554 tools/comm_bench.py --short
run f_tp${NG:-4}_c2_ipc 400 $TR --master-port 29557 bench.py --gpus ${NG:-4} --steps 64 --warmup 8
B200_TP_NVLS=1 run f_tp${NG:-4}_c2_nvls 400 $TR --master-port 29558 bench.py --gpus ${NG:-4} --steps 64 --warmup 8
