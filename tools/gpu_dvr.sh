#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_model.py tests/test_gpu_dvr.py -q -m gpu > gpurun_out/t_dvr.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/t_dvr.log | cut -c1-500
