#!/bin/bash
# end-of-round check: whole GPU suite, smoke, the default bench command, its ncu launch list, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_benchref.json 2>/dev/null; cut -c1-300 gpurun_out/r02_benchref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench100m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/launch_list.py gpurun_out/r02_launches_bench100m.csv 61 | tail -4
