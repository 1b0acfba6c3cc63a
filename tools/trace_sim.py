"""Cycle accounting of the scoring kernel's MMA-issuing thread and one epilogue warp (ERN_TRACE_PTR profiling aid).

Per persistent unit the kernel adds to 8 counters: [0] cycles of the MMA thread's work loop, [1] gallery tiles it
issued, [2] cycles waiting for operand stages (TMA), [3] cycles waiting for a free accumulator buffer (epilogue),
[4] cycles waiting for the query tile, [5] cycles of epilogue warp 0's work loop, [6] of which waiting for MMAs,
[7] of which inside segment compactions.  Slots 2..7 are only filled by a -DERN_SIM_TRACE_WAITS build
(ab_libs/libern_tracewaits.so, ERN_B200_LIB): reading the clock around every wait slows the issue loop by ~8 %, so the
product build only keeps [0] and [1].
A 256 x 256 x 640 tile is 40 MMAs of 128 cycles at the tensor-core floor: 5120 cycles.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=4096)
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=640)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--order", default="random")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    trace = torch.zeros(2 * 148 * 8, dtype=torch.int64, device=dev)
    os.environ["ERN_TRACE_PTR"] = hex(trace.data_ptr())
    from fashionern_aaai2024_b200 import ops
    from bench import make_gallery
    gal = make_gallery(args.n, args.dim, dev, 1, args.order)
    gen = torch.Generator(device=dev).manual_seed(2)
    pred = torch.nn.functional.normalize(torch.randn(args.q, args.dim, generator=gen, device=dev), dim=-1).bfloat16()
    ops.sim_topk(pred, gal, args.k, check_overflow=False)
    torch.cuda.synchronize()
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        ops.sim_topk(pred, gal, args.k, check_overflow=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    tail = trace.view(2, 148, 8)[1].cpu().double()
    raw = trace.view(2, 148, 8)[0].cpu()
    n_comp = (raw[:, 4] >> 40).double().sum().item()
    raw[:, 4] &= (1 << 40) - 1
    t = raw.double()
    t = t[t[:, 1] > 0]
    tiles = t[:, 1].sum().item()
    per_tile = (t[:, 0].sum() / tiles).item()
    floor = 128.0 * 4 * (args.dim // 64)
    out = {"q": args.q, "n": args.n, "order": args.order, "ms": ms, "tflops": 2.0 * args.q * args.n * args.dim / ms / 1e9,
           "units": int(t.shape[0]), "tiles_per_unit_min_max": [int(t[:, 1].min().item() / args.iters), int(t[:, 1].max().item() / args.iters)],
           "mma_thread_cycles_per_tile": per_tile, "floor_cycles_per_tile": floor, "issue_efficiency": floor / per_tile,
           "wait_operands_per_tile": (t[:, 2].sum() / tiles).item(), "wait_accumulator_per_tile": (t[:, 3].sum() / tiles).item(),
           "wait_query_tile_per_tile": (t[:, 4].sum() / tiles).item(),
           "issue_blocked_per_tile": per_tile - ((t[:, 2] + t[:, 3] + t[:, 4]).sum() / tiles).item(),
           "epilogue_cycles_per_tile": (t[:, 5].sum() / tiles).item(), "epilogue_wait_mma_per_tile": (t[:, 6].sum() / tiles).item(),
           "compaction_cycles_per_tile_warp0": (t[:, 7].sum() / tiles).item(), "compactions_per_1k_tiles_warp0": 1e3 * n_comp / tiles,
           "warp0_tiles_busy_over_4k_8k_16k_per_1k_tiles": [1e3 * tail[:, j].sum().item() / tiles for j in range(3)],
           "warp0_max_tile_busy_cycles": tail[:, 3].max().item(),
           "mean_sm_clock_mhz_in_kernel": (t[:, 0].mean().item() / args.iters) / (ms * 1e-3) / 1e6,
           "unit_cycles_min_max": [t[:, 0].min().item() / args.iters, t[:, 0].max().item() / args.iters]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
