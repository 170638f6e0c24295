mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_imagenet_gpu.py tests/test_ops_gpu.py -x -q -k "moe or gemm" 2>&1 | tail -8
timeout 600 python tools/config_timing.py 5 2>/dev/null | tail -1
NDIT_MOE_GROUPED=0 timeout 600 python tools/config_timing.py 5 2>/dev/null | tail -1
timeout 300 python tools/gemm_bench.py 2>&1 | tail -8
