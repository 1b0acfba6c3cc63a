#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/t_all.log | cut -c1-400
python tools/quick_bench.py --q 64 --n 4000000 --iters 3 | tail -1
python tools/quick_bench.py --q 256 --n 4000000 --iters 3 | tail -1
python tools/quick_bench.py --q 1024 --n 4000000 --iters 3 | tail -1
