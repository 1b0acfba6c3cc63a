#!/bin/bash
# end-of-round check: whole GPU suite, smoke, the default bench command, its ncu launch list, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_benchref.json 2>/dev/null; cut -c1-300 gpurun_out/r02_benchref.json
bash tools/gpu_launches100m.sh
