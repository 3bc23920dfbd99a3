python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench/sharded.py --what 3,4 --frames 48 2> gpurun_out/r2ac_n8.err > gpurun_out/r2ac_n8.out
grep "^{" gpurun_out/r2ac_n8.out > gpurun_out/r2ac_n8.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2ac_n8.jsonl'):
    d=json.loads(l)
    print(d['config'][:11], {k:d[k] for k in ('frames_per_s','ms_per_frame','checks','wait_timeouts','kernel_ms_per_frame_by_rank') if k in d})
PY
tail -2 gpurun_out/r2ac_n8.err
