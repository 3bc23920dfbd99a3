python bench/kernel_ab.py --vols xor,bonsai --layouts 4 --skips 1 --bricks 2,1,0 --batch 30 --launches 12 2>&1 | tee gpurun_out/ab_bricks2.log | cut -c1-150
for b in 0; do python bench/sharded.py --what 3,4 --frames 24 --brick $b 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('brick $b', d['config'][:11], {k:d[k] for k in ('frames_per_s','ms_per_frame','generate_s','bricks') if k in d}, d.get('checks'))"; done 2>&1 | tee gpurun_out/ab_config34_auto.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; tail -3 gpurun_out/r2z_tests.log
