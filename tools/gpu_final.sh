#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench100m.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench100m.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_dvr.py tests/test_gpu_combiner.py tests/test_gpu_visualsr.py -q -m gpu -x -k "golden" > gpurun_out/sanitizer_heads.log 2>&1
echo "sanitizer heads rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_heads.log | tail -2
