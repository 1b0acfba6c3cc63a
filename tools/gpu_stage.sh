#!/bin/bash
python -m pytest tests/test_gpu_sim.py tests/test_gpu_fuzz.py -q -m gpu 2>&1 | tail -2
for i in 1 2; do python tools/quick_bench.py --n 10000000 --iters 5 | tail -1 | cut -c1-150; done
python tools/quick_bench.py --n 1000000 --iters 10 | tail -1 | cut -c1-150
python tools/quick_bench.py --q 64 --n 8000000 --iters 5 | tail -1 | cut -c1-150
