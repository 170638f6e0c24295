#!/bin/bash
# Run each GPU test group in its own process with a timeout so one hung kernel cannot eat the whole
# gpurun call.  Logs go to gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 5 $to python -m pytest -x -q -m gpu "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
  if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/$name.log | head -20 | tee -a gpurun_out/summary.txt; fi
}
: > gpurun_out/summary.txt
for g in "$@"; do
  case $g in
    gemm)   run gemm_id 60 tests/test_ops_gpu.py -k "identity"; run gemm 120 tests/test_ops_gpu.py -k "gemm_store"; run swiglu 120 tests/test_ops_gpu.py -k "swiglu";;
    rows)   run lnrope 90 tests/test_ops_gpu.py -k "ln_rope"; run resid 90 tests/test_ops_gpu.py -k "resid";;
    attnref) run attnref 120 tests/test_ops_gpu.py -k "attention and refkernel";;
    attn)   run attn 90 tests/test_ops_gpu.py -k "attention and tcgen05";;
    attn48) run attn48ref 120 tests/test_imagenet_gpu.py -k "head_dim_48 and refkernel"; run attn48 120 tests/test_imagenet_gpu.py -k "head_dim_48 and tcgen05";;
    imagenet) run imagenet_ref 300 tests/test_imagenet_gpu.py -k "forward_with_cfg and refkernel"; run imagenet 300 tests/test_imagenet_gpu.py -k "forward_with_cfg and tcgen05"; run imagenet_traj 120 tests/test_imagenet_gpu.py -k "sampler";;
    flag)   run attn96ref 120 tests/test_flagdit_gpu.py -k "head_dim_96 and refkernel"; run attn96 120 tests/test_flagdit_gpu.py -k "head_dim_96 and tcgen05"; run flag_ref 200 tests/test_flagdit_gpu.py -k "forward_with_cfg and refkernel"; run flag 200 tests/test_flagdit_gpu.py -k "forward_with_cfg and tcgen05"; run flag_traj 120 tests/test_flagdit_gpu.py -k "sampler"; run flag5b 400 tests/test_flagdit_gpu.py -k "flagship";;
    moe)    run moe 300 tests/test_imagenet_gpu.py -k "moe";;
    model)  run model 600 tests/test_model_gpu.py;;
    all)    run all 1200 tests;;
  esac
done
cat gpurun_out/summary.txt
