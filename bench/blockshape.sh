#!/bin/bash
# exploration: frame time of the bench kernel and M0 for several block shapes
for b in 8x8 8x4 8x16 4x8 16x4 16x8 8x32 32x4; do
  echo "== VKRT_BLOCK=$b"
  VKRT_BLOCK=$b python bench/run_variant.py 1 3 1 12 | tail -1
  VKRT_BLOCK=$b python bench/run_variant.py 0 2 1 12 | tail -1
  VKRT_BLOCK=$b python bench/run_variant.py 1 3 1 12 1920 1080 bonsai | tail -1
done
