#!/bin/bash
# A/B on one box: the previous build (build/ab/libvokselis_rt_old.so, made from git HEAD) vs this build with its
# switches; every variant twice, interleaved. usage: bench/ab2.sh [frames]
F=${1:-38}
run() {  # label, env...
  local label=$1; shift
  echo "== $label"
  env "$@" python bench/run_variant.py 1 3 1 $F | grep -o "median.*"
  env "$@" python bench/run_variant.py 0 2 1 $F | grep -o "median.*"
  env "$@" python bench/run_variant.py 1 3 1 $F 1920 1080 bonsai | grep -o "median.*"
}
for rep in 1 2; do
run "old build" VKRT_LIB=$PWD/build/ab/libvokselis_rt_old.so
run "new, defaults (closed form >= 64)"  VKRT_X=0
run "new, no cull" VKRT_CULL=0
run "new, closed-form leaps >= 4" VKRT_LEAP_CLOSED_MIN=4
run "new, closed-form never" VKRT_LEAP_CLOSED_MIN=1000000
done
