mkdir -p gpurun_out
L=lumina_t2x_b200
./tools/exp_mt_bench.bin > gpurun_out/r2d_exp_mt_bench.log 2>&1
cat gpurun_out/r2d_exp_mt_bench.log
cp $L/libndit_b200.so $L/libndit_b200_g3.so
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:attention_hr -s 5 -c 1 -f -o gpurun_out/prof_attn_hr python tools/attn_bench.py $L/libndit_b200_g3.so:3 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
