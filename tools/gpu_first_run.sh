#!/bin/bash
# First-contact GPU run: every stage in its own process (a trapped kernel poisons its CUDA context only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n 15 gpurun_out/$name.log; }
run t_fp32      python -m pytest tests/test_gpu_sim.py -q -m gpu -k "fp32_validation or fewer_rows" -x
run t_single    python -m pytest tests/test_gpu_sim.py -q -m gpu -k "single_cta" 
run t_pair      python -m pytest tests/test_gpu_sim.py -q -m gpu -k "cta_pair"
run t_simrest   python -m pytest tests/test_gpu_sim.py -q -m gpu -k "not fp32_validation and not fewer_rows and not single_cta and not cta_pair and not large_gallery"
run t_comb      python -m pytest tests/test_gpu_combiner.py -q -m gpu
run t_metrics   python -m pytest tests/test_gpu_metrics.py -q -m gpu
run t_large     python -m pytest tests/test_gpu_sim.py -q -m gpu -k "large_gallery"
run b_pair      python tools/quick_bench.py
ERN_FORCE_SINGLE_CTA=1 run b_single python tools/quick_bench.py
run b_pair10m   python tools/quick_bench.py --n 10000000 --iters 3
