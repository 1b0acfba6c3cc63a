#!/bin/bash
# refresh of the single-GPU evidence with the final code: head timing, bench lines, query-size sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_combiner.py -q -m gpu -x 2>&1 | tail -1
timeout 300 python tools/bench_head_small.py > gpurun_out/r02_bench_head_small.jsonl 2>/dev/null; cut -c1-120 gpurun_out/r02_bench_head_small.jsonl
python tools/stamp_head.py > gpurun_out/r02_stamp_head.jsonl 2>/dev/null; cut -c1-160 gpurun_out/r02_stamp_head.jsonl
bash tools/gpu_r2_bench.sh > gpurun_out/bench_set.log 2>&1
timeout 600 python tools/sweep_queries.py --qs 1,32,128,129,192,256,384,512,1024,2048,4096 2>&1 | cut -c1-150
cp gpurun_out/sweep_queries_d640.json gpurun_out/r02_sweep_queries_10m.json
