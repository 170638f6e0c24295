mkdir -p gpurun_out
bash tools/gpu_diag.sh all
python tools/config_timing.py 1 5 > gpurun_out/r2b_config_timing.jsonl 2> gpurun_out/r2b_config_timing.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
bash tools/gpu_ncu.sh ln_rope_qk4 resid_rms_mod4 final_norm gemv_rows patch_embed transpose_v
tail -3 gpurun_out/all.log; cat gpurun_out/r2b_config_timing.jsonl; tail -2 gpurun_out/r2b_config_timing.err
