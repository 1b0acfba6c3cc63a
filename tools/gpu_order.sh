#!/bin/bash
# ordered-gallery correctness + cost: parity tests, then random vs clustered timings and traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_order.py tests/test_gpu_fuzz.py -q -m gpu -x > gpurun_out/t_sim.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/t_sim.log | cut -c1-300
for o in random clustered; do
  timeout 400 python tools/quick_bench.py --n 10000000 --iters 20 --order $o 2>&1 | tail -1 | cut -c1-400
done
for o in random clustered; do
  ERN_B200_LIB=ab_libs/libern_tracewaits.so timeout 300 python tools/trace_sim.py --n 10000000 --iters 10 --order $o 2>&1 | tail -1
done
for o in random clustered; do
  timeout 400 python tools/quick_bench.py --n 100000000 --iters 4 --order $o 2>&1 | tail -1 | cut -c1-400
done
