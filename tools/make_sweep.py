"""Collect the bench lines of the BASELINE.json configs[4] sweep (1M / 10M / 100M rows x 1 / 2 / 4 / 8 GPUs) from
profiles/ into profiles/r02_sweep.json (one compact row per run + scaling efficiency per gallery size)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
FILES = {(100_000_000, 1): "r02_bench100m.json", (10_000_000, 1): "r02_bench10m.json", (1_000_000, 1): "r02_bench1m.json"}
for rows in (1_000_000, 10_000_000, 100_000_000):
    for n in (2, 4, 8):
        FILES[(rows, n)] = f"r02_bench_{rows}_n{n}.json"

out = []
for (rows, n), f in sorted(FILES.items()):
    path = os.path.join(P, f)
    if not os.path.exists(path):
        continue
    d = json.loads(open(path).read().strip().splitlines()[-1])
    pc = d.get("parity_check") or {}
    out.append({"gallery_rows": rows, "n_gpus": n, "queries_per_s": d["value"], "e2e_queries_per_s": d["e2e"]["value"],
                "ms_per_step": d["ms_per_step"], "roofline_frac_sustained": d["roofline"]["frac"],
                "tflops_per_gpu": d["roofline"]["achieved"], "gpu_launches_per_step": d["gpu_launches"] / d["steps"],
                "sm_mhz": (d.get("clocks") or {}).get("sm_mhz"), "parity_check_ok": pc.get("ok"),
                "parity_missed_rows": pc.get("missed_rows"), "parity_max_rank_gap": pc.get("max_rank_gap"),
                "status_ok": d.get("status_ok"), "file": "profiles/" + f})
base = {r["gallery_rows"]: r["queries_per_s"] for r in out if r["n_gpus"] == 1}
for r in out:
    if r["gallery_rows"] in base:
        r["speedup_vs_1gpu"] = r["queries_per_s"] / base[r["gallery_rows"]]
        r["scaling_efficiency"] = r["speedup_vs_1gpu"] / r["n_gpus"]
json.dump(out, open(os.path.join(P, "r02_sweep.json"), "w"), indent=1)
for r in out:
    print(f"{r['gallery_rows']:>11} rows x {r['n_gpus']} GPU: {r['queries_per_s']:>12.0f} q/s  e2e {r['e2e_queries_per_s']:>12.0f}  "
          f"{r['ms_per_step']:8.3f} ms/step  frac {r['roofline_frac_sustained']:.3f}  eff {r.get('scaling_efficiency', float('nan')):.3f}  parity {r['parity_check_ok']}")
