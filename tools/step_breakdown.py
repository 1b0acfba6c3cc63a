"""Where does a small-shard step go?  (BASELINE.json configs[4], the 1M-row gallery on 8 GPUs = 125k rows per rank.)
torchrun --nproc-per-node N tools/step_breakdown.py --rows-per-rank 125000
Times, per step: host enqueue time (perf_counter, no sync), GPU time of the whole step (events), and the GPU time of
each part run on its own (fusion head, local scoring + top-k, exchange barrier + merge, recall)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fashionern_aaai2024_b200 import ops, sharded, synthetic as syn  # noqa: E402
from fashionern_aaai2024_b200.combiner import CombinerSimple  # noqa: E402
from bench import make_gallery  # noqa: E402


def ev_time(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / iters
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, host * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-rank", type=int, default=125000)
    ap.add_argument("--q", type=int, default=4096)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Q, D, K = args.q, 640, 100
    rows = args.rows_per_rank
    gallery = make_gallery(rows, D, dev, 5000 + rank, "random")
    class_of = torch.arange(rows * world, dtype=torch.int32, device=dev)
    head = CombinerSimple(D, 4 * D, 8 * D).to(dev).eval()
    head.load_state_dict(syn.combiner_state(7, D))
    img, txt = syn.features(8, Q, D).to(dev), syn.features(9, Q, D).to(dev)
    tgt = torch.zeros(Q, dtype=torch.int32, device=dev)
    with torch.no_grad():
        _, qb = head(img, txt, want_bf16=True)
    ex = "p2p" if world > 1 else "nccl"

    def step():
        with torch.no_grad():
            _, q = head(img, txt, want_bf16=True)
        _, ids, _, _ = sharded.sharded_topk(q, gallery, K, rank * rows, check_overflow=False, exchange=ex)
        ops.recall_at_k(ids, class_of, tgt, (1, 10, 50, 100))

    def only_head():
        with torch.no_grad():
            head(img, txt, want_bf16=True)

    out = {"world": world, "rows_per_rank": rows, "queries": Q}
    out["step_gpu_ms"], out["step_host_enqueue_ms"] = ev_time(step)
    out["head_gpu_ms"], out["head_host_ms"] = ev_time(only_head)
    out["sim_local_gpu_ms"], out["sim_local_host_ms"] = ev_time(lambda: ops.sim_topk(qb, gallery, K, want_keys=True, check_overflow=False))
    out["sharded_topk_gpu_ms"], out["sharded_topk_host_ms"] = ev_time(lambda: sharded.sharded_topk(qb, gallery, K, rank * rows, check_overflow=False, exchange=ex))
    ids = sharded.sharded_topk(qb, gallery, K, rank * rows, check_overflow=False, exchange=ex)[1]
    out["recall_gpu_ms"], out["recall_host_ms"] = ev_time(lambda: ops.recall_at_k(ids, class_of, tgt, (1, 10, 50, 100)))
    l0 = ops.launch_counter.n
    step()
    out["launches_per_step"] = ops.launch_counter.n - l0
    out["tensor_floor_ms"] = 2.0 * Q * rows * D / 1.4e15 * 1e3
    # the same step replayed from a CUDA graph (single GPU only: the symmetric-memory barrier is not captured here)
    if world == 1:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        out["step_cuda_graph_gpu_ms"], _ = ev_time(g.replay)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
