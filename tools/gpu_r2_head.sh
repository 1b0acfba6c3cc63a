#!/bin/bash
# fused small-batch head: parity tests of the in-tree build and of the variant builds in ab_libs/, then A/B timing
mkdir -p gpurun_out
for lib in "" ab_libs/libern_ring220.so; do
  echo "== tests with ${lib:-in-tree}"
  ERN_B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_combiner.py -q -m gpu 2>&1 | tail -2
done
for lib in ab_libs/libern_base.so ab_libs/libern_parred.so "" ab_libs/libern_ring220.so; do
  ERN_B200_LIB=$lib timeout 200 python tools/bench_head_small.py --cases 640:1,640:16,640:32,640:64,512:32,512:64 2>/dev/null
done | tee gpurun_out/r02_head_micro_ab.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['lib'], d['dim'], d['rows'], round(d['gpu_us_per_forward_cold_l2'],2))"
timeout 120 python tools/stamp_head.py 2>/dev/null | tee gpurun_out/r02_stamp_head_v2.jsonl
ERN_B200_LIB=ab_libs/libern_ring220.so timeout 120 python tools/stamp_head.py 2>/dev/null | tee gpurun_out/r02_stamp_head_v2_ring220.jsonl
