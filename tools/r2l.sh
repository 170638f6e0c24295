mkdir -p gpurun_out
L=lumina_t2x_b200
timeout 400 python tools/attn_bench.py $L/libndit_b200.so:1 $L/libndit_b200_t1.so:1 $L/libndit_b200_t2.so:1 $L/libndit_b200_t3.so:1 $L/libndit_b200_t4.so:1 $L/libndit_b200_t5.so:1 $L/libndit_b200_t3t.so:1 > gpurun_out/r2l_attn_bench.log 2>&1
cat gpurun_out/r2l_attn_bench.log
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "moe_token_gate" 2>&1 | tail -3
