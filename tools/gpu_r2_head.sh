#!/bin/bash
# fused small-batch head: parity tests of the in-tree build, then A/B timing against the variant builds in ab_libs/
# (ab_libs/ is git-ignored: a variant is the in-tree objects with ern_combiner_small.cu recompiled under a -D switch,
#  as tools/build_trace_lib.sh does for the scoring kernel; the switches of the measured variants lived only in the
#  commits named in DESIGN.md 4.3a / profiles/r02_notes.md)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_combiner.py -q -m gpu 2>&1 | tail -1
for lib in "" ab_libs/libern_lean2.so; do
  ERN_B200_LIB=$lib timeout 100 python tools/bench_head_small.py --cases 640:1,640:16,640:32,640:64,512:32,512:64 2>/dev/null
done | tee gpurun_out/r02_head_align_ab.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['lib'], d['dim'], d['rows'], round(d['gpu_us_per_forward_cold_l2'],2))"
