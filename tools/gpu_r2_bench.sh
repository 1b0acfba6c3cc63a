#!/bin/bash
# round 2: bench.py with the in-run parity check, ordered gallery, dataset configs
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 1500 gpurun_out/$name.json; echo; tail -n 3 gpurun_out/$name.err; }
run r02_bench10m --gallery-rows 10000000 --steps 5 --warmup 3
run r02_bench10m_clustered --gallery-rows 10000000 --gallery-order clustered --steps 5 --warmup 3 --no-cpu-baseline
run r02_bench_fiq --config fiq --steps 10
run r02_bench_cirr --config cirr --steps 10
run r02_bench_f200k --config f200k --steps 5
run r02_bench1m --gallery-rows 1000000 --steps 10 --warmup 3 --no-cpu-baseline
run r02_bench100m --steps 5 --warmup 3
run r02_bench100m_clustered --gallery-order clustered --steps 5 --warmup 3 --no-cpu-baseline
