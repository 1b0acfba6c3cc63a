#!/bin/bash
# last call of round 2: refresh the small lines with the final library (10M / 1M rows, the small-batch head table)
mkdir -p gpurun_out
timeout 200 python bench.py --gallery-rows 10000000 --no-cpu-baseline > gpurun_out/r02_bench10m_final.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench10m_final.json
timeout 100 python bench.py --gallery-rows 1000000 --no-cpu-baseline > gpurun_out/r02_bench1m_final.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench1m_final.json
timeout 100 python tools/bench_head_small.py 2>/dev/null | tee gpurun_out/r02_bench_head_small_final.jsonl | cut -c1-140
