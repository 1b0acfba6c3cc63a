#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_1m.csv python tools/quick_bench.py --iters 1 > gpurun_out/prof1.log 2>&1
tail -3 gpurun_out/prof1.log
