"""Device-side timing of the fusion head and VisualSR (secondary kernels of the path)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern  # noqa: E402
from fashionern_aaai2024_b200 import synthetic as syn  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    out = []
    for dim in (640, 512):
        head = ern.CombinerSimple(dim, 4 * dim, 8 * dim)
        head.load_state_dict(syn.combiner_state(1, dim))
        head = head.to(dev).eval()
        sr = ern.VisualSR(dim)
        sr.load_state_dict(syn.visualsr_state(2, dim))
        sr = sr.to(dev).eval()
        for rows in (32, 4096, 32768, 131072):
            a = torch.randn(rows, dim, device=dev)
            b = torch.randn(rows, dim, device=dev)
            with torch.no_grad():
                ms = timeit(lambda: head(a, b, want_bf16=True))
            flop = rows * (144.0 * dim * dim + 16 * dim)
            out.append({"op": "CombinerSimple", "dim": dim, "rows": rows, "ms": ms, "tflops": flop / ms / 1e9})
            if rows <= 32768:
                x = torch.randn(rows, 13, dim, device=dev)
                with torch.no_grad():
                    ms = timeit(lambda: sr(x))
                flop = rows * 14 * 2.0 * dim * dim
                out.append({"op": "VisualSR", "dim": dim, "rows": rows, "ms": ms, "tflops": flop / ms / 1e9,
                            "gbps_in": rows * 13 * dim * 4 / ms / 1e6})
    for dim in (640, 512):
        dvr = ern.DVR_module(dim)
        dvr.load_state_dict(syn.dvr_full_state(3, dim))
        dvr = dvr.to(dev).eval()
        for rows in (32, 512, 4096):
            pt, tk = torch.randn(rows, 13, dim, device=dev), torch.randn(rows, 77, dim, device=dev)
            a, b = torch.randn(rows, dim, device=dev), torch.randn(rows, dim, device=dev)
            with torch.no_grad():
                ms_enc = timeit(lambda: dvr.encode(pt, tk), iters=5)
                ms_all = timeit(lambda: dvr(pt, tk, a, b), iters=5)
            L, I = 91, 3072
            flop = rows * (2 * (L * (8.0 * dim * dim + 4.0 * dim * I) + 4.0 * L * L * dim) + 13 * 8.0 * dim * dim)
            out.append({"op": "DVR.encode", "dim": dim, "rows": rows, "ms": ms_enc, "tflops": flop / ms_enc / 1e9,
                        "queries_per_s": rows / ms_enc * 1e3, "ms_full_forward": ms_all})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
