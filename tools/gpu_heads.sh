#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_combiner.py tests/test_gpu_visualsr.py tests/test_gpu_dvr.py tests/test_gpu_full_model.py tests/test_gpu_metrics.py -q -m gpu > gpurun_out/t_heads.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/t_heads.log | cut -c1-300
python tools/bench_heads.py 2>&1 | grep -E "rows\": (4096|32768|512)," | cut -c1-220
