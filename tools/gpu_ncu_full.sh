#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sim_topk_tc -s 4 -c 1 -f -o gpurun_out/sim_full python tools/quick_bench.py --iters 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
