#!/bin/bash
# launch list of the DEFAULT bench command (100M rows) + sanitizer pass over the training-criterion tests
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ern|simtc|combiner|select|recall|gemm" --csv \
  --log-file gpurun_out/launches_bench100m.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench100m_under_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 1 gpurun_out/bench100m_under_ncu.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_loss.py -q -m gpu -x -k "golden or ragged" > gpurun_out/sanitizer_loss.log 2>&1
echo "sanitizer loss rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_loss.log | tail -2
