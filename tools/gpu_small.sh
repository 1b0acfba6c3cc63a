#!/bin/bash
for q in 32 64 128 256 512; do python tools/quick_bench.py --q $q --n 8000000 --iters 5 | tail -1 | cut -c1-160; done
python tools/quick_bench.py --n 100000000 --iters 2 | tail -1 | cut -c1-160
