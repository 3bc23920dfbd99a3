#!/bin/bash
for v in "1 0 1" "0 2 1" "0 0 1" "1 0 0"; do
  echo "== variant $v   new:"
  python bench/run_variant.py $v 12 | tail -2
  echo "   old:"
  VKRT_LIB=$PWD/build/ab/libvokselis_rt_old.so python bench/run_variant.py $v 12 | tail -2
done
