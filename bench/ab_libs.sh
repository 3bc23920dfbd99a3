#!/bin/bash
# A/B of whole library builds on the bench workload: build/ab/lib*.so (VKRT_LIB selects the library rt.py loads).
# usage: bench/ab_libs.sh A B C ...   -> gpurun_out/ab_libs.log
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    echo "== lib$v rep $rep"
    VKRT_LIB=$PWD/build/ab/lib$v.so python bench/kernel_ab.py --vols xor,bonsai --layouts 4 --skips 1 --batch 30 --launches 12
  done
done 2>&1 | tee gpurun_out/ab_libs.log
