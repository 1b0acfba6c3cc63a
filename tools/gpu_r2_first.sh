#!/bin/bash
# round 2, first GPU contact of the self-compacting candidate store: parity tests, then quick timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_order.py tests/test_gpu_fuzz.py -q -m gpu -x > gpurun_out/t_sim.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/t_sim.log | cut -c1-300
for n in 1000000 10000000; do
  for g in 8 16; do
    timeout 300 python tools/quick_bench.py --n $n --growth $g 2>&1 | tail -1
  done
done
for g in 8 16; do
  timeout 400 python tools/quick_bench.py --n 100000000 --growth $g --iters 3 2>&1 | tail -1
done
for t in 16 32 128; do
  ERN_TILES_PER_ITEM=$t timeout 400 python tools/quick_bench.py --n 100000000 --growth 8 --iters 3 2>&1 | tail -1
done
