#!/bin/bash
# A/B in the batched (issue-bound) regime: 8 frames per launch, 20 launches, mean launch ms
run() { echo "== $*"; env "$@" python bench/run_batch.py 8 20 | python -c "
import sys,re
s=sys.stdin.read(); v=[float(x) for x in re.findall(r'[0-9]+\.[0-9]+', s)]
print('mean launch ms %.4f -> %.0f frames/s' % (sum(v[2:])/len(v[2:]), 8e3/(sum(v[2:])/len(v[2:]))))"; }
for rep in 1 2; do
run VKRT_LEAP_CLOSED_MIN=64
run VKRT_LEAP_CLOSED_MIN=2
run VKRT_LEAP_CLOSED_MIN=4
run VKRT_LEAP_CLOSED_MIN=8
run VKRT_LEAP_CLOSED_MIN=16
run VKRT_LEAP_CLOSED_MIN=1000000
run VKRT_BLOCK=8x8
run VKRT_BLOCK=8x32
run VKRT_BLOCK=16x8
run VKRT_CULL=0
run VKRT_BBOX=0
done
