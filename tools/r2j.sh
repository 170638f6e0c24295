mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "sde" 2>&1 | tail -12
