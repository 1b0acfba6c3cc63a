#!/bin/bash
for pf in 0 2 4 8; do
  echo "pf=$pf"
  ERN_PREFETCH_TILES=$pf python tools/quick_bench.py --n 10000000 --iters 5 | tail -1
done
for pf in 0 4; do
  ERN_PREFETCH_TILES=$pf python tools/quick_bench.py --n 1000000 --iters 10 | tail -1
  ERN_PREFETCH_TILES=$pf python tools/quick_bench.py --n 60000000 --iters 2 | tail -1
done
python -m pytest tests/test_gpu_sim.py -q -m gpu 2>&1 | tail -1
