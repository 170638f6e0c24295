mkdir -p gpurun_out
timeout 600 python tools/attn_error.py > gpurun_out/r2i_attn_error.log 2>&1; grep -v "^  " gpurun_out/r2i_attn_error.log | tail -28
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "taps or plain or list" 2>&1 | tail -4
