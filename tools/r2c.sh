mkdir -p gpurun_out
L=lumina_t2x_b200
timeout 300 python tools/attn_bench.py $L/libndit_b200_hrt.so:3 $L/libndit_b200_hrtnt.so:3 > gpurun_out/r2c_attn_bench.log 2>&1
cat gpurun_out/r2c_attn_bench.log
