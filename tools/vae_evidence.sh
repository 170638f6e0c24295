#!/bin/bash
# VAE-decode end: timing vs the stock path, ncu launch list of one decode, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 300 python tools/vae_bench.py 128 10 > gpurun_out/vae_bench.json 2> gpurun_out/vae_bench.err; cat gpurun_out/vae_bench.json
timeout 300 python tools/vae_bench.py 256 5 --no-stock 2>/dev/null | cut -c1-260
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/vae_launches.csv python tools/vae_bench.py 128 1 --no-stock > gpurun_out/vae_ncu.log 2>&1; tail -1 gpurun_out/vae_ncu.log
# 3x3 conv 256 -> 256 at 1024^2 (BN 256, shared A tile) is gemm launch 33 of a decode, the 128-channel convs follow (34: 256 -> 128, 35: shortcut, 36..: 128 -> 128)
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 33 -c 1 -f -o gpurun_out/vae_prof_conv256 python tools/vae_bench.py 128 1 --no-stock > gpurun_out/vae_ncu_conv256.log 2>&1; tail -1 gpurun_out/vae_ncu_conv256.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 36 -c 1 -f -o gpurun_out/vae_prof_conv128 python tools/vae_bench.py 128 1 --no-stock > gpurun_out/vae_ncu_conv128.log 2>&1; tail -1 gpurun_out/vae_ncu_conv128.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:vae_gn_apply -s 26 -c 1 -f -o gpurun_out/vae_prof_gn_apply python tools/vae_bench.py 128 1 --no-stock > gpurun_out/vae_ncu_gn_apply.log 2>&1; tail -1 gpurun_out/vae_ncu_gn_apply.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:vae_gn_stats -s 26 -c 1 -f -o gpurun_out/vae_prof_gn_stats python tools/vae_bench.py 128 1 --no-stock > gpurun_out/vae_ncu_gn_stats.log 2>&1; tail -1 gpurun_out/vae_ncu_gn_stats.log
ls -la gpurun_out | grep vae_
