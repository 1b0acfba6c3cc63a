"""Row-sharded gallery over the GPUs of one box (north_star item 4).

The reference scores on a single device (run/test/test_fiq.py:138); this is new functionality.
One process per GPU.  Gallery rows are block-partitioned (rank r owns rows
[r*ceil(N/S), (r+1)*ceil(N/S)) and their global ids); queries and target ids are replicated.
Per query batch every rank

  1. runs the streaming top-k over ITS rows                       (ops.sim_topk, ids already global)
  2. exchanges the [Q,k] uint64 candidate keys (8 B per candidate; 3.3 MB per rank at Q=4096, k=100):
       exchange="p2p"  (default when symmetric memory is available): FUSED into the last selection launch --
                       the kernel that produces the keys stores them straight into every peer's gathered buffer
                       over NVLink (peer pointers into torch symmetric memory), then one device-side barrier;
       exchange="nccl" : ``all_gather_into_tensor`` of the locally written keys (also the gloo path in CPU tests)
  3. k-way merges the S sorted lists on the device                 (ops.topk_merge; ties -> lower global id)

so every rank ends with the identical global top-k.  There is no other data-path collective.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from ._lib import MODE_BF16, RANK_SIMILARITY


def exchange_in_use(requested: str) -> str:
    """'p2p' unless setting up symmetric memory failed (then every call silently uses 'nccl')."""
    return "nccl" if (requested == "p2p" and _p2p_error is not None) else requested


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


class _PeerBuffers:
    """Two symmetric-memory gathered buffers [S, Q, k] (double buffered so that a fast rank can start the next
    batch's stores while a slow rank still merges the previous one) + their peer-pointer tables."""

    def __init__(self, nq: int, k: int, device, group):
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.bufs, self.handles = [], []
        gname = (group or dist.group.WORLD).group_name
        for _ in range(2):
            t = symm.empty((self.world, nq, k), dtype=torch.int64, device=device)
            self.handles.append(symm.rendezvous(t, gname))
            self.bufs.append(t)
        self.turn = 0


_peer_cache = {}
_p2p_error = None      # set when symmetric memory could not be set up: callers then use the NCCL exchange


def _peer_buffers(nq, k, device, group) -> _PeerBuffers:
    key = (nq, k, device.index, id(group))
    if key not in _peer_cache:
        _peer_cache[key] = _PeerBuffers(nq, k, device, group)
    return _peer_cache[key]


def sharded_topk(queries: torch.Tensor, local_gallery: torch.Tensor, k: int, id_offset: int, *,
                 mode: int = MODE_BF16, rank_by: int = RANK_SIMILARITY, exclude_ids: Optional[torch.Tensor] = None,
                 group: Optional[dist.ProcessGroup] = None, check_overflow: bool = True, exchange: str = "nccl"):
    """Global top-k of every (replicated) query over a row-sharded gallery.
    Returns ``(values [Q,k], global ids [Q,k], keys [Q,k], status int32[4])``; the first three are identical on
    every rank, ``status`` is this rank's overflow report (see ``ops.sim_topk``)."""
    global _p2p_error
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    pb = None
    if multi and exchange == "p2p" and _p2p_error is None:
        try:
            pb = _peer_buffers(queries.shape[0], k, queries.device, group)
        except Exception as e:  # noqa: BLE001  (no P2P / symmetric-memory support on this box: NCCL path instead)
            _p2p_error = repr(e)
            import warnings
            warnings.warn(f"peer-memory candidate exchange unavailable ({_p2p_error}); using the NCCL all-gather")
    if pb is not None:
        buf, hdl = pb.bufs[pb.turn], pb.handles[pb.turn]
        pb.turn ^= 1
        status = ops.sim_topk_exchange(queries, local_gallery, k, hdl.buffer_ptrs_dev, pb.world, hdl.rank, mode=mode,
                                       rank_by=rank_by, exclude_ids=exclude_ids, id_offset=id_offset)
        if check_overflow and int(status[0].item()) != 0:
            raise ops.ErnError(f"ern_sim_topk_exchange reported an inconsistent candidate store ({status.tolist()})")
        hdl.barrier(channel=0)                                      # every rank's stores have landed everywhere
        return (*ops.topk_merge(buf, k), status)
    _, _, keys, status = ops.sim_topk(queries, local_gallery, k, mode=mode, rank_by=rank_by, exclude_ids=exclude_ids,
                                 id_offset=id_offset, want_keys=True, check_overflow=check_overflow)
    if not multi:
        return (*ops.topk_merge(keys.unsqueeze(0), k), status)
    gathered = exchange_candidates(keys, group)
    return (*ops.topk_merge(gathered, k), status)


def sharded_recall(queries: torch.Tensor, local_gallery: torch.Tensor, id_offset: int, class_of: torch.Tensor,
                   target_class: torch.Tensor, ks, *, mode: int = MODE_BF16, rank_by: int = RANK_SIMILARITY,
                   exclude_ids: Optional[torch.Tensor] = None, group=None, exchange: str = "nccl"):
    """Recall@K hit counts over a row-sharded gallery: global top-max(ks) (identical on every rank) followed by the
    id-membership kernel against the replicated ``class_of`` table (int32 [N_total]).  Returns (counts, ranks, ids)."""
    _, ids, _, _ = sharded_topk(queries, local_gallery, int(max(ks)), id_offset, mode=mode, rank_by=rank_by,
                                exclude_ids=exclude_ids, group=group, exchange=exchange)
    counts, ranks = ops.recall_at_k(ids, class_of, target_class, ks)
    return counts, ranks, ids


def sharded_cirr_subset(queries: torch.Tensor, local_gallery: torch.Tensor, id_offset: int, members: torch.Tensor,
                        reference_ids: torch.Tensor, target_ids: torch.Tensor, ks=(1, 2, 3), *, rank_by: int = 1,
                        group=None):
    """CIRR subset recall when the group members' rows are spread over the shards: each rank scores
    the members it owns, one all-reduce(SUM) of the [Q, 6] matrix assembles them, every rank ranks identically."""
    scores = ops.gather_scores(queries, local_gallery, members, id_offset)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(scores, op=dist.ReduceOp.SUM, group=group)
    return ops.cirr_subset_from_scores(scores, members, reference_ids, target_ids, ks, rank_by=rank_by)
