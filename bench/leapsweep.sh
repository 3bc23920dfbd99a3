#!/bin/bash
for v in 4 8 16 24 48 1000000; do
  echo "== VKRT_LEAP_CLOSED_MIN=$v"
  VKRT_LEAP_CLOSED_MIN=$v python bench/run_variant.py 1 3 1 14 | tail -1
  VKRT_LEAP_CLOSED_MIN=$v python bench/run_variant.py 0 2 1 14 | tail -1
  VKRT_LEAP_CLOSED_MIN=$v python bench/run_variant.py 1 3 1 14 1920 1080 bonsai | tail -1
  VKRT_LEAP_CLOSED_MIN=$v python bench/configs.py --config 4 --frames 8 --checks 0 2>&1 | grep config | python -c "import json,sys; print('config4', json.loads(sys.stdin.readline())['ms_per_frame'])"
done
