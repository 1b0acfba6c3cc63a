#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n 5 gpurun_out/$name.log | cut -c1-600; }
run t_sim    python -m pytest tests/test_gpu_sim.py tests/test_gpu_metrics.py -q -m gpu
run b_pair   python tools/quick_bench.py
run b_pair10m python tools/quick_bench.py --n 10000000 --iters 3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ern|simtc|combiner" -c 400 --csv --log-file gpurun_out/launches_bench10m.csv python bench.py --gallery-rows 10000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
