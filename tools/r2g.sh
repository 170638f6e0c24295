mkdir -p gpurun_out
L=lumina_t2x_b200
timeout 300 python tools/attn_bench.py $L/libndit_b200.so:1 $L/libndit_b200_bbs.so:1 $L/libndit_b200_bbst.so:1 > gpurun_out/r2g_attn_bench.log 2>&1
cat gpurun_out/r2g_attn_bench.log
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "attention and tcgen05 and not gen" 2>&1 | tail -3
MINE='regex:(gemm2?_bf16|attention_|resid_rms|ln_rope|transpose_v|gemv_rows|final_layer|final_norm|moe_|patch_embed|unpatchify|cond_prepare|rope_table|ln_rows|rms_rows|fill_ones|axpy)'
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 1500 --csv --log-file gpurun_out/r2f_launches.csv python tools/one_forward.py 2 > gpurun_out/r2f_ncu_launches.log 2>&1
tail -2 gpurun_out/r2f_ncu_launches.log
