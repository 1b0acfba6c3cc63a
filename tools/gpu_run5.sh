#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n 4 gpurun_out/$name.log | cut -c1-3000; }
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench10m python bench.py --gallery-rows 10000000 --steps 5 --warmup 3
run bench100m python bench.py
run benchref python bench.py --impl reference --steps 2 --warmup 1
