#!/bin/bash
# Multi-GPU suite for N GPUs (run under gpurun --gpus N): correctness checks, bench.py, configs 3/4 (sort-first), config 5 (sort-last)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29701 tests/mgpu_sortfirst_check.py 2>&1 | grep -E "MISMATCH|timeouts = [1-9]|Error" | head -5; echo "sortfirst check rc=${PIPESTATUS[0]}"
$TR --master-port 29702 tests/mgpu_sortlast_check.py 2>&1 | grep -E "sort-last|Error" | head -5
$TR --master-port 29703 bench.py --gpus $N --steps 360 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err | cut -c1-300; python bench/print_bench.py gpurun_out/bench_n$N.json
$TR --master-port 29704 bench/configs.py --frames 12 2>&1 | grep '^{"config"' > gpurun_out/configs34_n$N.jsonl; python -c "
import json
for l in open('gpurun_out/configs34_n$N.jsonl'):
    d=json.loads(l); print(d['config'][:9], 'N=$N', d['ms_per_frame'], 'ms', d['checks_at_full_size'])"
$TR --master-port 29705 bench/config5.py --edge 4096 --frames 4 2>&1 | grep '^{"config"' > gpurun_out/config5_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/config5_n$N.json')); print('config5 N=$N', d['ms_per_frame'], 'ms', d['per_frame_ms'])"
