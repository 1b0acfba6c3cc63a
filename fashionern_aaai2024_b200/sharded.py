"""Row-sharded gallery over the GPUs of one box (north_star item 4).

The reference scores on a single device (run/test/test_fiq.py:138); this is new functionality.
One process per GPU.  Gallery rows are block-partitioned (rank r owns rows
[r*ceil(N/S), (r+1)*ceil(N/S)) and their global ids); queries and target ids are replicated.
Per query batch every rank

  1. runs the streaming top-k over ITS rows                       (ops.sim_topk, ids already global)
  2. all-gathers the [Q,k] uint64 candidate keys over NCCL/NVLink  (8 B per candidate; 3.3 MB per rank at
     Q=4096, k=100 -- latency-, not bandwidth-bound on NVSwitch)
  3. k-way merges the S sorted lists on the device                 (ops.topk_merge; ties -> lower global id)

so every rank ends with the identical global top-k.  There is no other data-path collective.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from ._lib import MODE_BF16, RANK_SIMILARITY


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [begin, end) owned by ``rank``: equal blocks of ceil(N/S), the last ones possibly short/empty."""
    per = (n_rows + world_size - 1) // world_size
    begin = min(rank * per, n_rows)
    return begin, min(begin + per, n_rows)


def exchange_candidates(keys: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank candidate keys [Q,k] (int64 view of the uint64 wire format) -> [S,Q,k].
    Works on any backend (NCCL on the GPU box, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    q = keys.shape[0]
    out = torch.empty((world * q,) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, keys.contiguous(), group=group)   # concatenates along dim 0, in rank order
    return out.view((world, q) + tuple(keys.shape[1:]))


def sharded_topk(queries: torch.Tensor, local_gallery: torch.Tensor, k: int, id_offset: int, *,
                 mode: int = MODE_BF16, rank_by: int = RANK_SIMILARITY, exclude_ids: Optional[torch.Tensor] = None,
                 group: Optional[dist.ProcessGroup] = None, check_overflow: bool = True):
    """Global top-k of every (replicated) query over a row-sharded gallery.
    Returns ``(values [Q,k], global ids [Q,k], keys [Q,k], status int32[4])``; the first three are identical on
    every rank, ``status`` is this rank's overflow report (see ``ops.sim_topk``)."""
    _, _, keys, status = ops.sim_topk(queries, local_gallery, k, mode=mode, rank_by=rank_by, exclude_ids=exclude_ids,
                                 id_offset=id_offset, want_keys=True, check_overflow=check_overflow)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return (*ops.topk_merge(keys.unsqueeze(0), k), status)
    gathered = exchange_candidates(keys, group)
    return (*ops.topk_merge(gathered, k), status)
