mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
