#!/bin/bash
# round-2 evidence run: fuzz soak, cycle traces (product build + wait-split build), query-size sweep, L2 prefetch probe
mkdir -p gpurun_out
ERN_FUZZ_CASES=500 ERN_FUZZ_SEED=77 timeout 1200 python -m pytest tests/test_gpu_fuzz.py -q -m gpu -x 2>&1 | tail -2
for n in 10000000 100000000; do
  it=10; [ $n = 100000000 ] && it=3
  timeout 300 python tools/trace_sim.py --n $n --iters $it > gpurun_out/r02_trace_${n}.json 2>/dev/null; cut -c1-330 gpurun_out/r02_trace_${n}.json
  ERN_B200_LIB=ab_libs/libern_tracewaits.so timeout 300 python tools/trace_sim.py --n $n --iters $it > gpurun_out/r02_trace_waits_${n}.json 2>/dev/null
done
ERN_B200_LIB=ab_libs/libern_tracewaits.so timeout 300 python tools/trace_sim.py --n 10000000 --iters 10 --order clustered > gpurun_out/r02_trace_waits_10000000_clustered.json 2>/dev/null
timeout 600 python tools/sweep_queries.py --qs 1,32,128,129,192,256,384,512,1024,2048,4096 2>&1 | cut -c1-200
cp gpurun_out/sweep_queries_d640.json gpurun_out/r02_sweep_queries_10m.json
for pf in 1 2 3; do for q in 128 256; do ERN_PREFETCH_TILES=$pf timeout 200 python tools/quick_bench.py --q $q --n 10000000 --iters 20 | cut -c1-110 | sed "s/^/pf=$pf /"; done; done
