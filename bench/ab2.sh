#!/bin/bash
F=${1:-38}
run() {  # label, env...
  local label=$1; shift
  echo "== $label"
  env "$@" python bench/run_variant.py 1 3 1 $F 1920 1080 bonsai | grep -o "median.*\|stats.*"
  env "$@" python bench/run_variant.py 0 2 1 $F | grep -o "median.*\|stats.*"
  env "$@" python bench/run_variant.py 1 3 1 $F | grep -o "median.*\|stats.*"
}
for rep in 1 2; do
run "no occupied-bounds clip" VKRT_BBOX=0
run "clip, closed-form entry leap"  VKRT_BBOX=1
run "clip, no entry leap"  VKRT_BBOX=2
run "clip, replayed entry leap"  VKRT_BBOX=4
done
