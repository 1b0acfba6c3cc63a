#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sim_topk_tc -s 5 -c 1 -f -o gpurun_out/sim_full_10m python tools/quick_bench.py --n 10000000 --iters 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file gpurun_out/launches_bench10m.csv python bench.py --gallery-rows 10000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
