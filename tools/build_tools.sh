#!/bin/bash
# Build the development tools (not part of the product library) for sm_100a into build/.  They travel to the GPU box
# with the repo snapshot; none of them is loaded by the tests, smoke() or bench.py.
set -eu
cd "$(dirname "$0")/.."
mkdir -p build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
F="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr"
$NVCC $F -o build/umma_probe tools/umma_probe.cu &
$NVCC $F -o build/umma_rate tools/umma_rate.cu &
$NVCC $F -o build/tma_issue_rate tools/tma_issue_rate.cu &
$NVCC $F -DB200_GEMM_TRACE -o build/gemm_trace tools/gemm_trace.cu &
$NVCC $F -o build/gemm_bench tools/gemm_trace.cu &
$NVCC $F -DB200_GEMM_TOOL -o build/gemm_sweep tools/gemm_sweep.cu &
wait
ls -la build/umma_probe build/umma_rate build/tma_issue_rate build/gemm_trace build/gemm_bench build/gemm_sweep
