#!/bin/bash
# A/B in one GPU session over the builds in ab_libs/ (and the in-tree one = IN_TREE)
mkdir -p gpurun_out
LIBS="${LIBS:-ab_libs/libern_r01.so ab_libs/libern_NO_BOTH.so IN_TREE}"
for n in 10000000 100000000; do
  it=20; [ $n = 100000000 ] && it=4
  for lib in $LIBS; do
    [ $lib = IN_TREE ] && lib=""
    ERN_B200_LIB=$lib timeout 400 python tools/quick_bench.py --n $n --iters $it 2>&1 | tail -1 | cut -c1-400
  done
done
timeout 400 python tools/quick_bench.py --n 10000000 --iters 20 --order clustered 2>&1 | tail -1 | cut -c1-400
timeout 400 python tools/quick_bench.py --n 100000000 --iters 4 --order clustered 2>&1 | tail -1 | cut -c1-400
timeout 300 python tools/trace_sim.py --n 10000000 --iters 10 2>&1 | tail -1
timeout 300 python tools/trace_sim.py --n 100000000 --iters 3 2>&1 | tail -1
