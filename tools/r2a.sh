mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm" 2>&1 | tail -8 > gpurun_out/r2a_gemm_tests.log
timeout 1500 python -m pytest tests/test_reference_gpu.py -q -s 2>&1 | tail -60 > gpurun_out/r2a_ref_tests.log
python tools/gemm_bench.py > gpurun_out/r2a_gemm_bench.log 2>&1
python tools/attn_vs_fa2.py > gpurun_out/r2a_attn_vs_fa2.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
NDIT_RESID4=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-stock-cuda-baseline > gpurun_out/r2a_bench_resid4.json 2> /dev/null
tail -3 gpurun_out/r2a_gemm_tests.log; tail -5 gpurun_out/r2a_ref_tests.log
