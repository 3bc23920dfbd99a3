bench/ab_libs.sh D H > /dev/null; grep -E "==|fps" gpurun_out/ab_libs.log | cut -c1-130
python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_tests.log 2>&1; tail -3 gpurun_out/r2aa_tests.log
