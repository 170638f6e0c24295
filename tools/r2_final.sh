#!/bin/bash
# round-2 final evidence pass: full GPU suite, smoke, default bench, compositional timing, ncu of the region-masked attention kernel
mkdir -p gpurun_out
rm -f gpurun_out/reference_parity.jsonl
timeout -k 5 900 python -m pytest tests -m gpu -q -s --durations=15 2>&1 | grep -v "^$" > gpurun_out/r2h_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2h_pytest.log | tail -15
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench.json
timeout -k 5 200 python tools/config_timing.py 6 > gpurun_out/r2h_config_timing_6.jsonl 2>&1; tail -2 gpurun_out/r2h_config_timing_6.jsonl
timeout -k 5 240 ncu --set full --clock-control none --import-source on -k regex:attention_fused_kernel -s 30 -c 1 -f -o gpurun_out/r2h_prof_attention_region python tools/config_timing.py 6 > gpurun_out/r2h_ncu_region.log 2>&1
tail -2 gpurun_out/r2h_ncu_region.log
ls -la gpurun_out | tail -12
