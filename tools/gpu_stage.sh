#!/bin/bash
python -m pytest tests/test_gpu_sim.py tests/test_gpu_fuzz.py -q -m gpu 2>&1 | tail -2
python tools/quick_bench.py --n 10000000 --iters 5 --dim 640 | tail -1 | cut -c1-150
python tools/quick_bench.py --n 10000000 --iters 5 --dim 512 | tail -1 | cut -c1-150
python tools/quick_bench.py --n 10000000 --iters 5 --dim 256 | tail -1 | cut -c1-150
python tools/quick_bench.py --q 64 --n 8000000 --iters 5 --dim 512 | tail -1 | cut -c1-150
