mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2m_tests.log; tail -3 gpurun_out/r2m_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["clocks"], {k:(round(v["ms"],1) if isinstance(v,dict) else round(v,3)) for k,v in d["kernels"].items()}, d["gpu_launches"], d["roofline"]["frac"], d.get("stock_cuda_baseline",{}).get("engine_speedup"), d.get("cpu_baseline"))
PY
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; tail -c 600 gpurun_out/r2m_bench_ref.json
