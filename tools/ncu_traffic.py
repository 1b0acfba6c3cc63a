"""ncu --set full raw CSV of one scoring launch -> an entry of profiles/sim_topk_traffic.json (read by bench.py for
roofline.traffic).   python tools/ncu_traffic.py <raw.csv> <launch_rows> <queries> <dim> [--print]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, launch_rows, queries, dim = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rows = list(csv.reader(open(path)))
    hdr, units, val = rows[0], rows[1], rows[2]
    get = lambda name: next((float(val[i].replace(",", "")), units[i]) for i, h in enumerate(hdr) if h == name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd, ru = get("dram__bytes_read.sum")
    wr, wu = get("dram__bytes_write.sum")
    dur, du = get("gpu__time_duration.sum")
    dur_s = dur * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}[du]
    e = {"kernel": "ern::simtc::sim_topk_tc_kernel<true, 0>", "queries": queries, "dim": dim, "launch_rows": launch_rows,
         "dram_bytes_read": rd * scale[ru], "dram_bytes_write": wr * scale[wu],
         "algorithmic_bytes": launch_rows * dim * 2.0 + queries * dim * 2.0,
         "duration_s": dur_s, "tflops_in_capture": 2.0 * queries * launch_rows * dim / dur_s / 1e12,
         "tensor_pipe_active_pct_of_elapsed": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")[0],
         "lts_hit_rate_pct": get("lts__t_sector_hit_rate.pct")[0],
         "source": "profiles/" + os.path.basename(path)}
    e["dram_read_over_algorithmic"] = e["dram_bytes_read"] / e["algorithmic_bytes"]
    print(json.dumps(e, indent=1))
    if "--print" in sys.argv:
        return
    out = os.path.join(ROOT, "profiles", "sim_topk_traffic.json")
    entries = json.load(open(out)) if os.path.exists(out) else []
    entries = [x for x in entries if not (x["queries"] == queries and x["dim"] == dim and x["launch_rows"] == launch_rows)] + [e]
    json.dump(entries, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
