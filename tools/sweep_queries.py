"""Roofline sweep of ern_sim_topk over the query-batch size (10M x D gallery, k = 100): achieved TFLOP/s and gallery
GB/s beside the roofline time max(FLOP / bf16 peak, gallery bytes / HBM peak) from MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fashionern_aaai2024_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=640)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--qs", type=str, default="1,8,32,64,128,129,256,384,512,1024,2048,4096,8192,16384")
    args = ap.parse_args()
    peak_tf, peak_gbs = 1424.8, 6454.6
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        peak_tf = float(pk.get("bf16_tflops", peak_tf))         # burst figure: each size is timed alone, for < 0.5 s
        peak_gbs = float(pk.get("hbm_gbs", peak_gbs))
    except Exception:  # noqa: BLE001
        pass
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    gal = torch.empty(args.n, args.dim, dtype=torch.bfloat16, device=dev)
    step = 1 << 20
    for s in range(0, args.n, step):
        x = torch.randn(min(step, args.n - s), args.dim, generator=gen, device=dev)
        gal[s:s + step] = torch.nn.functional.normalize(x, dim=-1).bfloat16()
    out = []
    for q in [int(v) for v in args.qs.split(",")]:
        pred = torch.nn.functional.normalize(torch.randn(q, args.dim, generator=gen, device=dev), dim=-1).bfloat16()
        for _ in range(2):
            res = ops.sim_topk(pred, gal, args.k, check_overflow=False)
        torch.cuda.synchronize()
        iters = 5 if q >= 1024 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.sim_topk(pred, gal, args.k, check_overflow=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flop = 2.0 * q * args.n * args.dim
        gbytes = args.n * args.dim * 2 / 1e9
        ideal_ms = max(flop / (peak_tf * 1e12), gbytes / peak_gbs) * 1e3
        rec = {"q": q, "n": args.n, "dim": args.dim, "k": args.k, "ms": ms, "tflops": flop / ms / 1e9,
               "gallery_gbs": gbytes / ms * 1e3, "roofline_ms": ideal_ms, "frac_of_roofline": ideal_ms / ms,
               "bound": "tensor" if flop / (peak_tf * 1e12) > gbytes / peak_gbs else "hbm",
               "overflow": int(res[3][0].item())}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/sweep_queries_d{args.dim}.json", "w") as f:
        json.dump({"peak_tflops": peak_tf, "peak_gbs": peak_gbs, "rows": out}, f, indent=1)


if __name__ == "__main__":
    main()
