"""GPU time of one CombinerSimple.forward at the reference's query batch (32 rows; run/test/test_fiq.py:132), measured
without Python launch overhead: 50 forwards captured in one CUDA graph, replayed, CUDA events around the replay.
Floor: the 59 MB (D = 640) of bf16 weights once from HBM."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern  # noqa: E402
from fashionern_aaai2024_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="", help="dim:rows,dim:rows,... (default: the full table)")
    args = ap.parse_args()
    only = {tuple(int(x) for x in c.split(":")) for c in args.cases.split(",") if c}
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    for dim in (640, 512):
        head = ern.CombinerSimple(dim, 4 * dim, 8 * dim)
        head.load_state_dict(syn.combiner_state(1, dim))
        head = head.to(dev).eval()
        for rows in (1, 16, 32, 64, 65, 128):
            if only and (dim, rows) not in only:
                continue
            a, b = torch.randn(rows, dim, device=dev), torch.randn(rows, dim, device=dev)
            flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)     # 256 MB > L2, READ between forwards: the weights come from HBM and evict clean lines
            flush_sink = torch.zeros((), dtype=torch.float32, device=dev)
            with torch.no_grad():
                for _ in range(3):
                    head(a, b, want_bf16=True)
                torch.cuda.synchronize()
                reps = 20
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(reps):
                        flush_sink.copy_(flush.sum())
                        head(a, b, want_bf16=True)
                gf = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gf):
                    for _ in range(reps):
                        flush_sink.copy_(flush.sum())
            def t(graph):
                graph.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    graph.replay()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / 5 / reps
            us = (t(g) - t(gf)) * 1e3
            wbytes = 72.0 * dim * dim * 2
            print(json.dumps({"op": "CombinerSimple.forward", "lib": os.path.basename(os.environ.get("ERN_B200_LIB", "in-tree")), "dim": dim, "rows": rows, "gpu_us_per_forward_cold_l2": us,
                              "weight_mb": wbytes / 1e6, "weight_gbs": wbytes / us / 1e3,
                              "frac_of_hbm_peak": wbytes / us / 1e3 / peaks["hbm_gbs"],
                              "floor_us": wbytes / peaks["hbm_gbs"] / 1e3}))


if __name__ == "__main__":
    main()
