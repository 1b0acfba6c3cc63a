#!/bin/bash
mkdir -p gpurun_out
# memcheck + synccheck + racecheck on small shapes (the sanitizer is this repo's race detector)
# round 2: includes the self-pruning segments (ordered galleries), the warp-per-query selection and the fused small-batch head
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_sim.py tests/test_gpu_order.py -q -m gpu -x \
     -k "fp32_validation_mode_matches_oracle and 48-192 or single_cta_matches_oracle and 5-130 or cta_pair_matches_oracle and 129-257 or fewer_rows or duplicate_rows or split_cirr or (ascending and 513) or (ascending and 3-30000-64-100-0)" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_combiner.py tests/test_gpu_visualsr.py -q -m gpu -x -k "golden or (small_batches and 640 and (32 or 17 or 64))" > gpurun_out/sanitizer_heads.log 2>&1
echo "heads rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_heads.log | tail -2
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_combiner.py -q -m gpu -x -k "small_batches and 640 and 32" > gpurun_out/sanitizer_heads_race.log 2>&1
echo "heads racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_heads_race.log | tail -2
grep -h "Race reported\|hazard" gpurun_out/sanitizer_racecheck.log gpurun_out/sanitizer_heads_race.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -8
