#!/bin/bash
mkdir -p gpurun_out
# memcheck + synccheck + racecheck on small shapes (the sanitizer is this repo's race detector)
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_sim.py -q -m gpu -x \
     -k "fp32_validation_mode_matches_oracle and 48-192 or single_cta_matches_oracle and 5-130 or cta_pair_matches_oracle and 129-257 or fewer_rows or duplicate_rows or split_cirr" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_combiner.py tests/test_gpu_visualsr.py -q -m gpu -x -k "golden" > gpurun_out/sanitizer_heads.log 2>&1
echo "heads rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_heads.log | tail -2
