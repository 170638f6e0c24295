#!/bin/bash
# round-2 evidence pass: launch list of one forward, ncu --set full of every kernel class, racecheck of the mbarrier-heavy kernels
mkdir -p gpurun_out
MINE='regex:(gemm2?_bf16|attention_|resid_rms|ln_rope|transpose_v|gemv_rows|final_layer|final_norm|moe_|patch_embed|unpatchify|cond_prepare|rope_table|ln_rows|rms_rows|fill_ones|axpy|place_rows)'
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 700 --csv --log-file gpurun_out/r2f_launches.csv python tools/one_forward.py 2 > gpurun_out/r2f_ncu_launches.log 2>&1
for k in gemm2_bf16_tn_kernel gemm_bf16_tn_kernel attention_fused_kernel resid_rms_mod4 ln_rope_qk4 gemv_rows final_norm patch_embed unpatchify_cfg cond_prepare; do
  timeout -k 5 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2f_prof_$k python tools/one_forward.py 1 > gpurun_out/r2f_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2f_ncu_$k.log
done
NDIT_ATTN_GEN=3 timeout -k 5 240 ncu --set full --clock-control none --import-source on -k regex:attention_hr -s 3 -c 1 -f -o gpurun_out/r2f_prof_attention_hr python tools/one_forward.py 1 > gpurun_out/r2f_ncu_attention_hr.log 2>&1
tail -1 gpurun_out/r2f_ncu_attention_hr.log
# racecheck (shared-memory hazards; mbarrier / TMA / tcgen05 kernels): small operator tests only, it is slow
timeout -k 5 500 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_ops_gpu.py -x -q -k "(gemm_store or swiglu or attention) and not refkernel" > gpurun_out/r2f_racecheck.log 2>&1
tail -5 gpurun_out/r2f_racecheck.log
ls -la gpurun_out | head -40
