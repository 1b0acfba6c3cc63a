#!/bin/bash
# ncu launch list (gpu__time_duration) of this library's kernels inside the DEFAULT bench command, + per-kernel share of one step
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"sim_topk_tc_kernel|select_topk|init_state_kernel|recall_kernel|gemm_tc_kernel|finalize_kernel|cast_bf16|zero_i32|fused_head|l2norm" \
  -c 700 --csv --log-file gpurun_out/r02_launches_bench100m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections, json
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_bench100m.csv")) if len(r) > 10]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
data = rows[1:]
# one step of the default command = 6 fusion-head launches + 1 init + 53 x (scoring, warp select, block select) + merge + 2 recall = 169
last = data[-169:]
agg = collections.OrderedDict()
for r in last:
    k = r[ki][:64]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
out = {"command": "python bench.py --steps 2 --warmup 1 (100M rows, Q = 4096, k = 100)", "launches_in_step": len(last), "step_ms_under_ncu": tot / 1e6,
       "kernels": [{"kernel": k, "launches": v[0], "us": v[1] / 1e3, "share_pct": 100 * v[1] / tot} for k, v in agg.items()]}
json.dump(out, open("gpurun_out/r02_bench100m_step_shares.json", "w"), indent=1)
for k, v in agg.items():
    print(f"{v[0]:3d} x {k:66s} {v[1]/1e3:10.1f} us {100*v[1]/tot:6.2f} %")
print("launches profiled", len(data), "step ms", tot / 1e6)
PY
