for rep in 1 2; do for v in I J; do for b in 20 30; do
echo "== lib$v batch $b"; VKRT_LIB=$PWD/build/ab/lib$v.so python bench/kernel_ab.py --vols xor --layouts 4 --skips 1 --batch $b --launches 12 | cut -c1-120
done; done; done 2>&1 | tee gpurun_out/ab_blockcull.log
python -m pytest tests -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1; tail -2 gpurun_out/r3c_tests.log
