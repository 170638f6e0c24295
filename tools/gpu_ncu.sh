#!/bin/bash
# ncu captures of full-size forward_with_cfg calls (config 2).  Outputs under gpurun_out/.
#   launch list: every launch of OUR kernels (second forward of two), device time per launch
#   full sets  : one launch of each kernel named on the command line
mkdir -p gpurun_out
MINE='regex:(gemm2?_bf16|attention_|resid_rms|ln_rope|transpose_v|gemv_rows|final_layer|final_norm|moe_|patch_embed|unpatchify|cond_prepare|rope_table|ln_rows|rms_rows|fill_ones|axpy)'
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 700 --csv --log-file gpurun_out/launches.csv python tools/one_forward.py 2 > gpurun_out/ncu_launches.log 2>&1
for k in "$@"; do
  timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python tools/one_forward.py 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
