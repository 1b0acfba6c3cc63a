"""GPU time of one CombinerSimple.forward at the reference's query batch (32 rows; run/test/test_fiq.py:132), measured
without Python launch overhead: 50 forwards captured in one CUDA graph, replayed, CUDA events around the replay.
Floor: the 59 MB (D = 640) of bf16 weights once from HBM.
``--flags-sweep``: the same measurement for every ERN_HEAD_FLAGS value (scheduling choices of the fused kernel, see
csrc/ern_combiner_small.cu), outputs checked bit-identical against flags 0."""
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
    ap.add_argument("--flags-sweep", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    cases = [(dim, rows, None) for dim in (640, 512) for rows in (1, 16, 32, 64, 65, 128)]
    if args.flags_sweep:
        cases = [(dim, rows, fl) for dim, rows in ((640, 32), (640, 1), (640, 64), (512, 32)) for fl in range(8)]
    heads, want = {}, {}
    for dim, rows, flags in cases:
        if dim not in heads:
            head = ern.CombinerSimple(dim, 4 * dim, 8 * dim)
            head.load_state_dict(syn.combiner_state(1, dim))
            heads[dim] = head.to(dev).eval()
        head = heads[dim]
        if flags is None:
            os.environ.pop("ERN_HEAD_FLAGS", None)
        else:
            os.environ["ERN_HEAD_FLAGS"] = str(flags)           # (the library reads it on every call)
        for _ in (0,):
            gen = torch.Generator(device=dev).manual_seed(rows)
            a, b = torch.randn(rows, dim, device=dev, generator=gen), torch.randn(rows, dim, device=dev, generator=gen)
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
            same = None
            if flags is not None:
                with torch.no_grad():
                    out = head(a, b, want_bf16=True)[0].clone()
                same = bool(torch.equal(out, want.setdefault((dim, rows), out)))   # flags 0 comes first
            print(json.dumps({"op": "CombinerSimple.forward", "dim": dim, "rows": rows, "flags": flags,
                              "bit_identical_to_flags0": same, "gpu_us_per_forward_cold_l2": us,
                              "weight_mb": wbytes / 1e6, "weight_gbs": wbytes / us / 1e3,
                              "frac_of_hbm_peak": wbytes / us / 1e3 / peaks["hbm_gbs"],
                              "floor_us": wbytes / peaks["hbm_gbs"] / 1e3}))


if __name__ == "__main__":
    main()
