"""Quick device-side timing of ern_sim_topk (not the contract bench; see bench.py)."""
import argparse
import json
import sys
import os

import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fashionern_aaai2024_b200 import ops  # noqa: E402
from bench import ClockSampler, make_gallery  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=4096)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=640)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--growth", type=int, default=8)
    ap.add_argument("--order", default="random", choices=["random", "clustered"])
    ap.add_argument("--operands", default="bf16", choices=["bf16", "fp16"], help="16-bit operand format (ERN_DTYPE_*)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    gal = make_gallery(args.n, args.dim, dev, 1, args.order)
    pred = torch.nn.functional.normalize(torch.randn(args.q, args.dim, generator=gen, device=dev), dim=-1).bfloat16()
    if args.operands == "fp16":                 # same values (bf16 -> fp16 is exact here: unit-norm entries), other format
        blk = 1 << 20
        gal16 = torch.empty_like(gal, dtype=torch.float16)
        for s in range(0, args.n, blk):
            gal16[s:s + blk] = gal[s:s + blk].half()
        gal, pred = gal16, pred.half()
    for _ in range(2):
        out = ops.sim_topk(pred, gal, args.k, growth=args.growth, check_overflow=False)
    torch.cuda.synchronize()
    st = out[3].cpu().tolist()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        ops.sim_topk(pred, gal, args.k, growth=args.growth, check_overflow=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    clocks = sampler.stop(t0, time.time())
    flops = 2.0 * args.q * args.n * args.dim
    print(json.dumps({"q": args.q, "n": args.n, "dim": args.dim, "k": args.k, "ms": ms,
                      "tflops": flops / ms / 1e9, "qps": args.q / ms * 1e3, "status": st, "order": args.order, "operands": args.operands,
                      "sm_mhz": clocks.get("sm_mhz"), "power_w": clocks.get("power_w_max"), "lib": os.path.basename(os.environ.get("ERN_B200_LIB", "head")),
                      "single": os.environ.get("ERN_FORCE_SINGLE_CTA", "0")}))


if __name__ == "__main__":
    main()
