"""CPU: host-side logic of the row-sharded path, including a world_size-2 gloo run of the candidate
exchange (the CUDA merge itself is covered by tests/test_gpu_sim.py::test_shard_merge_equals_global)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fashionern_aaai2024_b200 import sharded
from fashionern_aaai2024_b200.metrics import factorize_names, percent
from fashionern_aaai2024_b200.ops import _phase_count


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 100, 1001, 100_000_000):
        for world in (1, 2, 4, 8):
            spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) == (n + world - 1) // world or n == 0


def test_factorize_and_percent():
    cls, table, counts = factorize_names(["a", "b", "a", "c"])
    assert cls.tolist() == [0, 1, 0, 2] and table == {"a": 0, "b": 1, "c": 2} and counts.tolist() == [2, 1, 1]
    assert percent(7, 2017) == 0.3470500698313117


def test_phase_schedule():
    # dense launch over the first 256 rows, then x8 gallery ranges (csrc/ern_capi.cu)
    assert _phase_count(100, 50, 8) == 1 and _phase_count(256, 50, 8) == 1
    assert _phase_count(257, 50, 8) == 2 and _phase_count(2048, 50, 8) == 2 and _phase_count(2049, 50, 8) == 3
    # 256, 2k, 16k, 131k, 1M, 8.4M, 67M, 100M: a launch is exact for any number of rows ...
    assert _phase_count(1_000_000, 100, 8) == 5 and _phase_count(10_000_000, 100, 8) == 7
    assert _phase_count(100_000_000, 100, 8) == 8 and _phase_count(100_000_000, 100, 16) == 6
    # ... but tensor-core launches are cut at 2M rows: keeps the query tiles of a super tile aligned (L2 sharing) and
    # the thresholds fresh on ordered galleries
    assert _phase_count(100_000_000, 100, 8, 1 << 21) == 5 + 1 + -(-(100_000_000 - (1 << 20) - (1 << 21)) // (1 << 21))
    assert _phase_count(10_000, 100, 1) == 1 + -(-(10_000 - 256) // (2048 - 100))
    # the fp32 validation kernel never covers more rows than a query has candidate slots (74 x 256 here)
    assert _phase_count(100_000, 100, 8, 74 * 256) == 3 + -(-(100_000 - 16384) // (74 * 256))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        keys = torch.randint(0, 2 ** 62, (q, k), generator=g, dtype=torch.int64).sort(dim=1, descending=True).values
        out = sharded.exchange_candidates(keys)
        ret[rank] = out.numpy()
    finally:
        dist.destroy_process_group()


def test_candidate_exchange_world2_gloo():
    world, q, k = 2, 5, 7
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), q, k, ret), nprocs=world, join=True)
        a, b = ret[0], ret[1]
    assert a.shape == (world, q, k) and np.array_equal(a, b)          # every rank sees the same lists
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        exp = torch.randint(0, 2 ** 62, (q, k), generator=g, dtype=torch.int64).sort(dim=1, descending=True).values
        assert np.array_equal(a[r], exp.numpy())                      # in rank order
