#!/bin/bash
# end of round 2: whole GPU suite, smoke, the default bench command, compute-sanitizer over the changed small-batch head
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_default.json
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_combiner.py -q -m gpu -x -k "small_batches and 640 and (32 or 17 or 64 or 1-)" > gpurun_out/sanitizer_heads_v2.log 2>&1
echo "heads memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_heads_v2.log | tail -2
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_combiner.py -q -m gpu -x -k "small_batches and 640 and (32 or 64)" > gpurun_out/sanitizer_heads_race_v2.log 2>&1
echo "heads racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_heads_race_v2.log | tail -2
grep -h "Race reported\|hazard" gpurun_out/sanitizer_heads_race_v2.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -6
