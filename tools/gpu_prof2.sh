#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2; do
  ERN_DEBUG_FLAGS=$f python tools/quick_bench.py --iters 3 2>&1 | tail -1
  ERN_DEBUG_FLAGS=$f ERN_FORCE_SINGLE_CTA=1 python tools/quick_bench.py --iters 3 2>&1 | tail -1
done
