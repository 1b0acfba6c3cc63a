#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n 6 gpurun_out/$name.log; }
run t_sim    python -m pytest tests/test_gpu_sim.py -q -m gpu
run b_pair   python tools/quick_bench.py
ERN_FORCE_SINGLE_CTA=1 run b_single python tools/quick_bench.py
run b_pair10m python tools/quick_bench.py --n 10000000 --iters 3
run b_pair100m python tools/quick_bench.py --n 100000000 --iters 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_1m.csv python tools/quick_bench.py --iters 1 > gpurun_out/prof1.log 2>&1
