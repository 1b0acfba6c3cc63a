#!/bin/bash
# late round 2: whole GPU suite (new fp16 / single-GPU exchange tests), exact-fraction census, smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/t_all.log | cut -c1-300
timeout 200 python tools/exact_frac_census.py > gpurun_out/r02_exact_frac_census.jsonl 2>&1; grep min_exact gpurun_out/r02_exact_frac_census.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_default.json
