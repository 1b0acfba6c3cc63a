#!/bin/bash
# round 2, late: fp16 operand format on the scoring path + scheduling flags of the fused small-batch head
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_combiner.py tests/test_gpu_sim.py tests/test_gpu_metrics.py -q -m gpu > gpurun_out/t_flags.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/t_flags.log | cut -c1-300
timeout 300 python tools/bench_head_small.py --flags-sweep > gpurun_out/r02_head_flags.jsonl 2> gpurun_out/r02_head_flags.err; echo "sweep rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_head_flags.jsonl"):
    d = json.loads(l)
    print(d["dim"], d["rows"], d["flags"], d["bit_identical_to_flags0"], round(d["gpu_us_per_forward_cold_l2"], 2))
PY
for op in bf16 fp16; do timeout 200 python tools/quick_bench.py --n 10000000 --operands $op --iters 10 2>&1 | tail -1; done | tee gpurun_out/r02_quick_fp16_vs_bf16.jsonl
