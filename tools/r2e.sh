mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2e_tests.log
tail -6 gpurun_out/r2e_tests.log
python tools/gemm_bench.py > gpurun_out/r2e_gemm_bench.log 2>&1; cat gpurun_out/r2e_gemm_bench.log
for g in 1 3; do
NDIT_ATTN_GEN=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-stock-cuda-baseline > gpurun_out/r2e_bench_gen$g.json 2> gpurun_out/r2e_bench_gen$g.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_bench_gen$g.json").read().strip().splitlines()[-1])
print("gen$g", d["value"], d["ms_per_step"], d["clocks"], {k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d["kernels"].items()}, d["gpu_launches"])
PY
done
NDIT_VT_EPI=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-stock-cuda-baseline > gpurun_out/r2e_bench_novtepi.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_bench_novtepi.json").read().strip().splitlines()[-1])
print("vt_epi=0", d["value"], d["ms_per_step"], {k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d["kernels"].items()}, d["gpu_launches"])
PY
