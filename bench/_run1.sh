python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3_bench_n8.out 2> gpurun_out/r3_bench_n8.err
grep "^{" gpurun_out/r3_bench_n8.out > gpurun_out/r3_bench_n8.json; python bench/print_bench.py gpurun_out/r3_bench_n8.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r3_bench_n4.out 2> gpurun_out/r3_bench_n4.err
grep "^{" gpurun_out/r3_bench_n4.out > gpurun_out/r3_bench_n4.json; python bench/print_bench.py gpurun_out/r3_bench_n4.json | cut -c1-300
