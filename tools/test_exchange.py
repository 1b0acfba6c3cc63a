"""torchrun --nproc-per-node N tools/test_exchange.py : fused peer-memory exchange == NCCL all-gather path,
bit for bit, over several back-to-back batches (exercises the double buffering) -- multi-GPU check."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fashionern_aaai2024_b200 import ops, sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n, dim, q, k = 3_000_001, 640, 700, 100
    begin, end = sharded.shard_bounds(n, world, rank)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    gal = torch.nn.functional.normalize(torch.randn(end - begin, dim, generator=gen, device=dev), dim=-1).bfloat16()
    ok = True
    for it in range(6):
        g = torch.Generator(device=dev).manual_seed(7 + it)          # same queries on every rank
        pred = torch.nn.functional.normalize(torch.randn(q, dim, generator=g, device=dev), dim=-1).bfloat16()
        if it % 2 == rank % 2:
            time.sleep(0.05)                                         # skew the ranks
        v1, i1, k1, s1 = sharded.sharded_topk(pred, gal, k, begin, exchange="nccl")
        v2, i2, k2, s2 = sharded.sharded_topk(pred, gal, k, begin, exchange="p2p")
        same = torch.equal(k1, k2) and torch.equal(i1, i2) and torch.equal(v1, v2)
        # every rank must hold the same global answer
        ref = k2.clone()
        dist.broadcast(ref, 0)
        same = same and torch.equal(ref, k2)
        ok = ok and same
    # the sharded answer must equal the answer of one GPU holding the whole gallery (rank 0 gathers the shards)
    parts = [torch.empty(sharded.shard_bounds(n, world, r)[1] - sharded.shard_bounds(n, world, r)[0], dim,
                         dtype=torch.bfloat16, device=dev) if rank == 0 else None for r in range(world)]
    if rank == 0:
        parts[0].copy_(gal)
        for r in range(1, world):
            dist.recv(parts[r], src=r)
        whole = torch.cat(parts)
        _, i_one, _, _ = ops.sim_topk(pred, whole, k)
        single_ok = torch.equal(i_one, i2)
        del whole
    else:
        dist.send(gal, dst=0)
        single_ok = True
    # timing of the two exchanges (small shard so that the exchange is visible)
    res = {}
    for ex in ("nccl", "p2p"):
        for _ in range(3):
            sharded.sharded_topk(pred, gal, k, begin, exchange=ex, check_overflow=False)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            sharded.sharded_topk(pred, gal, k, begin, exchange=ex, check_overflow=False)
        e1.record(); torch.cuda.synchronize()
        res[ex] = e0.elapsed_time(e1) / 20
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("EXCHANGE_OK" if int(flag.item()) == 1 else "EXCHANGE_MISMATCH", res)
        print("SINGLE_GPU_EQUAL" if single_ok else "SINGLE_GPU_MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
