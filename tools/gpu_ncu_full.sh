#!/bin/bash
# ncu --set full of a typical launch of the 100M-row step: the 6th scoring launch of a call covers gallery rows
# [1M, 3M) = 2097152 rows x 4096 queries (launches are cut at 2M rows).  Raw CSV -> gpurun_out/ (copy the summary into profiles/).
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:sim_topk_tc -s 5 -c 1 -o gpurun_out/r02_sim_topk_100m -f \
    python tools/quick_bench.py --n 100000000 --iters 1 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu_full.log
ncu -i gpurun_out/r02_sim_topk_100m.ncu-rep --page raw --csv > gpurun_out/r02_sim_topk_tc_100m_ncu_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_sim_topk_tc_100m_ncu_full_raw.csv")))
hdr, val = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__grid_size"]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want) or "tensor" in h and "pct" in h:
        print(h, rows[1][i], val[i])
PY
