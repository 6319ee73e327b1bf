#!/bin/bash
# A/B of decode attention variants (1 GPU): HEAD build, post-scaling with exact-int operands, post-scaling with fp16-subnormal
# operands, each with L2 prefetch off / 3 / 6.  The variant libraries (libb200tm_head.so, libb200tm_exactint.so) were one-off builds
# of decode_attention.cu linked against the other objects (nvcc ... -DB200_RAW_DENORMAL=0), loaded through B200_LIB_PATH; the
# experiment's outcome is in profiles/r02_attention_variants.md and the variants are not kept in the tree.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/lmdeploy_b200/lib
{
for lib in libb200tm_head.so libb200tm_exactint.so libb200tm.so; do
  for pf in 0 3 6; do
    [ $lib = libb200tm_head.so ] && [ $pf != 0 ] && continue
    B200_LIB_PATH=$L/$lib B200_ATTN_PREFETCH=$pf timeout 300 python tools/attn_ab.py 8 40 2>&1 | tail -n 1
  done
done
B200_LIB_PATH=$L/libb200tm_head.so timeout 300 python tools/attn_ab.py 4 40 2>&1 | tail -n 1
B200_LIB_PATH=$L/libb200tm.so timeout 300 python tools/attn_ab.py 4 40 2>&1 | tail -n 1
} | tee gpurun_out/attn_ab.log
