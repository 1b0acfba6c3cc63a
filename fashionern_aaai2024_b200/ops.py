"""Tensor-level wrappers over the C ABI: torch owns memory and streams, libern_b200 does the work.

Nothing here computes on the host or with torch ops -- torch is used for allocation, stream handles and
(in callers) ``torch.distributed``.  All functions require CUDA tensors on an sm_100-class device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import (DENSE_ROWS, DTYPE_BF16, DTYPE_F16, DTYPE_F32, MODE_BF16, MODE_FP32, QUERY_BATCH, RANK_REFERENCE,
                   RANK_SIMILARITY, SEG_CAP, SORT_CAP, NORM_OUT_F16, ErnError)

__all__ = ["l2norm_rows", "sim_topk", "sim_topk_exchange", "topk_merge", "recall_at_k", "cirr_subset_recall", "gather_scores",
           "cirr_subset_from_scores", "bbc_loss_forward", "bbc_loss_backward", "launch_counter"]


class _LaunchCounter:
    """Counts kernel launches issued through this module (bench.py reports it as ``gpu_launches``).
    The numbers mirror what the C ABI launches per call ('launch accounting')."""

    def __init__(self):
        self.n = 0

    def add(self, n: int):
        self.n += int(n)


launch_counter = _LaunchCounter()
_DTYPE_OF = {torch.float32: DTYPE_F32, torch.bfloat16: DTYPE_BF16, torch.float16: DTYPE_F16}


def _feature_dtype(q: torch.Tensor, g: torch.Tensor, mode: Optional[int] = None) -> int:
    """ERN_DTYPE_* of a (queries, gallery) pair: both float32 (fp32 validation mode) or both the same 16-bit type
    (tensor-core mode: bf16, or fp16 for 3 more mantissa bits at the same tcgen05 rate)."""
    if q.dtype != g.dtype or q.dtype not in _DTYPE_OF:
        raise ErnError(f"queries/gallery must both be float32, bfloat16 or float16, got {q.dtype} / {g.dtype}")
    if mode is not None and (mode == MODE_FP32) != (q.dtype == torch.float32):
        want = "float32" if mode == MODE_FP32 else "bfloat16 or float16"
        raise ErnError(f"mode {mode} takes {want} features, got queries {q.dtype} / gallery {g.dtype}")
    return _DTYPE_OF[q.dtype]


# rows per tensor-core scoring launch (mirrors launch_max_rows() in csrc/ern_capi.cu, incl. its environment override)
LAUNCH_MAX_ROWS = (lambda v: None if v <= 0 else v)(int(os.environ.get("ERN_LAUNCH_MAX_ROWS", str(1 << 21))))


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _rowmajor(t: torch.Tensor, what: str) -> torch.Tensor:
    L.require_cuda(t, what)
    if t.dim() != 2:
        raise ErnError(f"{what} must be 2-D, got shape {tuple(t.shape)}")
    if t.stride(1) != 1:
        t = t.contiguous()
    return t


def l2norm_rows(x: torch.Tensor, normalize: bool = True, want_f32: bool = True, want_bf16: bool = False,
                want_f16: bool = False) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """``F.normalize(x, dim=-1).float()`` (run/test/test_fiq.py:45) and/or its 16-bit rounding (bf16, or fp16 with
    ``want_f16``: saturating, for unit-norm features), on device.  Returns ``(fp32 | None, 16-bit | None)``."""
    x = _rowmajor(x, "x")
    if x.dtype != torch.float32:
        raise ErnError(f"l2norm_rows takes float32 features, got {x.dtype}")
    rows, dim = x.shape
    of = torch.empty((rows, dim), dtype=torch.float32, device=x.device) if want_f32 else None
    if want_bf16 and want_f16:
        raise ErnError("l2norm_rows emits one 16-bit copy: want_bf16 or want_f16, not both")
    ob = (torch.empty((rows, dim), dtype=torch.float16 if want_f16 else torch.bfloat16, device=x.device)
          if (want_bf16 or want_f16) else None)
    with torch.cuda.device(x.device):
        L.check(L.lib().ern_l2norm_rows(x.data_ptr(), rows, dim, x.stride(0),
                                        int(bool(normalize)) | (NORM_OUT_F16 if want_f16 else 0), L.ptr(of), dim,
                                        L.ptr(ob), dim, L.stream_ptr(x.device)))
    launch_counter.add(1 if rows else 0)
    return of, ob


def _phase_count(n_rows: int, k: int, growth: int, max_rows_per_launch: Optional[int] = None) -> int:
    """Number of scoring launches ern_sim_topk issues per query batch (mirrors the schedule in csrc/ern_capi.cu):
    rows [0,256) densely, then [b, growth*b) -- the fp32 validation kernel additionally caps a launch at the
    number of candidate slots a query owns (``max_rows_per_launch``)."""
    n, begin = 1, min(n_rows, DENSE_ROWS)
    while begin < n_rows:
        end = begin + (SORT_CAP - k) if growth == 1 else begin * growth
        if max_rows_per_launch is not None:
            end = min(end, begin + max_rows_per_launch)
        begin = end
        n += 1
    return n


def _sim_launches(nq: int, n_rows: int, k: int, growth: int, mode: int, device) -> int:
    """Kernel launches of one ern_sim_topk call: per query batch one state-init launch plus, per schedule step, a
    scoring launch and the selection (bf16 path: warp-per-query kernel + the block kernel for its leftovers)."""
    if nq == 0:
        return 0
    total = 0
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    for q0 in range(0, nq, QUERY_BATCH):
        bq = min(QUERY_BATCH, nq - q0)
        cap = LAUNCH_MAX_ROWS
        if mode == MODE_FP32:
            cap = (sms // 2 if bq > 128 else sms) * SEG_CAP
        total += 1 + (2 if mode == MODE_FP32 else 3) * _phase_count(n_rows, k, growth, cap)
    return total


def sim_topk(queries: torch.Tensor, gallery: torch.Tensor, k: int, *, mode: int = MODE_BF16,
             rank_by: int = RANK_SIMILARITY, exclude_ids: Optional[torch.Tensor] = None, id_offset: int = 0,
             growth: int = 8, want_keys: bool = False, check_overflow: bool = True):
    """Top-k gallery rows per query by cosine similarity; replaces ``1 - pred @ index.T`` + ``torch.argsort``
    (run/test/test_fiq.py:49-50) without materialising the [Q,N] matrix.

    Returns ``(values fp32 [Q,k], ids int32 [Q,k], keys uint64-as-int64 [Q,k] | None, status int32[4])``.
    The result is exact for any gallery order (candidate segments compact themselves inside the kernel), so there
    is no fallback path: ``status[0]`` can only be non-zero on an internal inconsistency, and
    ``check_overflow=True`` (one host sync) turns that into an ``ErnError``.
    """
    q = _rowmajor(queries, "queries")
    g = _rowmajor(gallery, "gallery")
    dtype = _feature_dtype(q, g, mode)
    if q.shape[1] != g.shape[1]:
        raise ErnError(f"feature dims differ: {q.shape[1]} vs {g.shape[1]}")
    if q.device != g.device:
        raise ErnError("queries and gallery must live on the same device")
    nq, dim = q.shape
    n_rows = g.shape[0]
    dev = q.device
    vals = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev) if want_keys else None
    status = torch.empty(4, dtype=torch.int32, device=dev)
    if exclude_ids is not None:
        L.require_cuda(exclude_ids, "exclude_ids")
        exclude_ids = exclude_ids.to(torch.int32).contiguous()
    lib = L.lib()
    with torch.cuda.device(dev):
        wsb = lib.ern_sim_topk_workspace_bytes(nq, dim, mode)
        ws = _workspace(wsb, dev)

        if nq == 0:
            status.zero_()
            return vals, ids, keys, status
        L.check(lib.ern_sim_topk(q.data_ptr(), nq, q.stride(0), g.data_ptr(), n_rows, g.stride(0), dim, dtype,
                                 int(id_offset), L.ptr(exclude_ids), int(k), mode, rank_by, growth,
                                 vals.data_ptr(), ids.data_ptr(), L.ptr(keys), status.data_ptr(),
                                 ws.data_ptr(), wsb, L.stream_ptr(dev)))
        launch_counter.add(_sim_launches(nq, n_rows, k, growth, mode, dev))
        if check_overflow and int(status[0].item()) != 0:
            raise ErnError(f"ern_sim_topk reported an inconsistent candidate store (status {status.tolist()})")
    return vals, ids, keys, status


def sim_topk_exchange(queries: torch.Tensor, gallery: torch.Tensor, k: int, peer_ptrs_dev: int, world: int, rank: int,
                      *, mode: int = MODE_BF16, rank_by: int = RANK_SIMILARITY,
                      exclude_ids: Optional[torch.Tensor] = None, id_offset: int = 0, growth: int = 8) -> torch.Tensor:
    """``sim_topk`` whose last launch writes this rank's [Q,k] candidate keys directly into slot ``rank`` of every
    rank's gathered buffer through peer pointers (``peer_ptrs_dev``: device address of an array of ``world``
    buffer pointers, e.g. ``torch.distributed._symmetric_memory`` ``buffer_ptrs_dev``).  Returns ``status``."""
    q = _rowmajor(queries, "queries")
    g = _rowmajor(gallery, "gallery")
    dtype = _feature_dtype(q, g, mode)
    nq, dim = q.shape
    dev = q.device
    status = torch.empty(4, dtype=torch.int32, device=dev)
    if exclude_ids is not None:
        exclude_ids = exclude_ids.to(torch.int32).contiguous()
    lib = L.lib()
    with torch.cuda.device(dev):
        wsb = lib.ern_sim_topk_workspace_bytes(nq, dim, mode)
        ws = _workspace(wsb, dev)
        L.check(lib.ern_sim_topk_exchange(q.data_ptr(), nq, q.stride(0), g.data_ptr(), g.shape[0], g.stride(0), dim,
                                          dtype, int(id_offset),
                                          L.ptr(exclude_ids), int(k), mode, rank_by, growth, peer_ptrs_dev, world,
                                          rank, status.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr(dev)))
        launch_counter.add(_sim_launches(nq, g.shape[0], k, growth, mode, dev))
    return status


def topk_merge(keys: torch.Tensor, k_out: int):
    """Merge per-shard sorted candidate lists ``keys`` [S, Q, k_in] (int64 view of the uint64 wire keys)
    into the global top-``k_out`` per query -- the device-side k-way merge after the NCCL all-gather."""
    L.require_cuda(keys, "keys")
    if keys.dim() != 3 or keys.dtype != torch.int64:
        raise ErnError("keys must be int64 [S, Q, k_in]")
    keys = keys.contiguous()
    s, nq, k_in = keys.shape
    dev = keys.device
    vals = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
    out_keys = torch.empty((nq, k_out), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ern_topk_merge(keys.data_ptr(), nq, s, k_in, nq * k_in, k_in, k_out, vals.data_ptr(),
                                       ids.data_ptr(), out_keys.data_ptr(), L.stream_ptr(dev)))
    launch_counter.add(1 if nq else 0)
    return vals, ids, out_keys


def _ks(ks: Sequence[int]):
    arr = (C.c_int32 * len(ks))(*[int(x) for x in ks])
    return arr, len(ks)


def recall_at_k(top_ids: torch.Tensor, class_of: torch.Tensor, target_class: torch.Tensor, ks: Sequence[int]):
    """Hit counts for Recall@K from id membership (run/test/test_fiq.py:51-60; any-hit for non-unique names,
    run/test/test_200k.py:52-60).  Returns ``(counts int32[len(ks)], first_hit_rank int32[Q])`` on device."""
    for t, n in ((top_ids, "top_ids"), (class_of, "class_of"), (target_class, "target_class")):
        L.require_cuda(t, n)
        if t.dtype != torch.int32:
            raise ErnError(f"{n} must be int32")
    top_ids, class_of, target_class = top_ids.contiguous(), class_of.contiguous(), target_class.contiguous()
    nq, k = top_ids.shape
    dev = top_ids.device
    counts = torch.empty(len(ks), dtype=torch.int32, device=dev)
    ranks = torch.empty(nq, dtype=torch.int32, device=dev)
    arr, nk = _ks(ks)
    with torch.cuda.device(dev):
        L.check(L.lib().ern_recall_at_k(top_ids.data_ptr(), nq, k, class_of.data_ptr(), class_of.numel(),
                                        target_class.data_ptr(), arr, nk, counts.data_ptr(), ranks.data_ptr(),
                                        L.stream_ptr(dev)))
    launch_counter.add(2 if nq else 1)
    return counts, ranks


def cirr_subset_recall(queries: torch.Tensor, gallery: torch.Tensor, members: torch.Tensor,
                       reference_ids: torch.Tensor, target_ids: torch.Tensor, ks: Sequence[int] = (1, 2, 3),
                       rank_by: int = RANK_REFERENCE):
    """CIRR subset recall (run/test/test_cirr.py:64-66,76-78).  ``members`` int32 [Q, m<=8] gallery rows.
    Returns ``(counts int32[len(ks)], rank int32[Q])``; rank -1 where the target is not a surviving member."""
    q = _rowmajor(queries, "queries")
    g = _rowmajor(gallery, "gallery")
    dtype = _feature_dtype(q, g)
    members = members.to(torch.int32).contiguous()
    reference_ids = reference_ids.to(torch.int32).contiguous()
    target_ids = target_ids.to(torch.int32).contiguous()
    nq, m = members.shape
    dev = q.device
    counts = torch.empty(len(ks), dtype=torch.int32, device=dev)
    ranks = torch.empty(nq, dtype=torch.int32, device=dev)
    arr, nk = _ks(ks)
    with torch.cuda.device(dev):
        L.check(L.lib().ern_cirr_subset_recall(q.data_ptr(), nq, q.stride(0), g.data_ptr(), g.shape[0], g.stride(0),
                                               q.shape[1], dtype, members.data_ptr(), m, reference_ids.data_ptr(),
                                               target_ids.data_ptr(), rank_by, arr, nk, counts.data_ptr(),
                                               ranks.data_ptr(), L.stream_ptr(dev)))
    launch_counter.add(2 if nq else 1)
    return counts, ranks


def gather_scores(queries: torch.Tensor, gallery: torch.Tensor, ids: torch.Tensor, id_offset: int = 0) -> torch.Tensor:
    """Similarity of query q with the gallery rows ``ids[q, :]`` (global ids) that live in this shard
    (rows ``[id_offset, id_offset + len(gallery))``); 0 for rows owned by another shard, so that a SUM over the
    shards gives the full [Q, m] matrix (CIRR group-member scores gathered from the owning shards)."""
    q = _rowmajor(queries, "queries")
    g = _rowmajor(gallery, "gallery")
    dtype = _feature_dtype(q, g)
    ids = ids.to(torch.int32).contiguous()
    nq, m = ids.shape
    out = torch.empty((nq, m), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        L.check(L.lib().ern_gather_scores(q.data_ptr(), nq, q.stride(0), g.data_ptr(), g.shape[0], g.stride(0), q.shape[1],
                                          dtype, int(id_offset),
                                          ids.data_ptr(), m, out.data_ptr(), L.stream_ptr(q.device)))
    launch_counter.add(1 if nq else 0)
    return out


def cirr_subset_from_scores(scores: torch.Tensor, members: torch.Tensor, reference_ids: torch.Tensor,
                            target_ids: torch.Tensor, ks: Sequence[int] = (1, 2, 3), rank_by: int = RANK_REFERENCE):
    """CIRR subset recall (run/test/test_cirr.py:64-66,76-78) from already gathered member similarities [Q, m]."""
    L.require_cuda(scores, "scores")
    scores = scores.float().contiguous()
    members = members.to(torch.int32).contiguous()
    reference_ids = reference_ids.to(torch.int32).contiguous()
    target_ids = target_ids.to(torch.int32).contiguous()
    nq, m = members.shape
    dev = scores.device
    counts = torch.empty(len(ks), dtype=torch.int32, device=dev)
    ranks = torch.empty(nq, dtype=torch.int32, device=dev)
    arr, nk = _ks(ks)
    with torch.cuda.device(dev):
        L.check(L.lib().ern_cirr_subset_from_scores(scores.data_ptr(), nq, members.data_ptr(), m, reference_ids.data_ptr(),
                                                    target_ids.data_ptr(), rank_by, arr, nk, counts.data_ptr(),
                                                    ranks.data_ptr(), L.stream_ptr(dev)))
    launch_counter.add(2 if nq else 1)
    return counts, ranks


def bbc_loss_forward(pred: torch.Tensor, tar: torch.Tensor, scale: float = 100.0, mode: int = MODE_BF16
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-batch classification loss (losses/loss.py:10-14): mean cross entropy of ``scale * pred @ tar.T`` against
    ``arange(B)``.  Returns (loss [1] fp32, row logsumexp [B] fp32)."""
    p = _rowmajor(pred, "pred")
    t = _rowmajor(tar, "tar")
    if p.dtype != torch.float32 or t.dtype != torch.float32 or p.shape != t.shape:
        raise ErnError("pred / tar must be float32 tensors of the same [B, D] shape")
    b, d = p.shape
    loss = torch.empty(1, dtype=torch.float32, device=p.device)
    lse = torch.empty(b, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        nbytes = L.lib().ern_bbc_loss_workspace_bytes(b, d, mode)
        ws = _workspace(nbytes, p.device)
        L.check(L.lib().ern_bbc_loss_forward(p.data_ptr(), p.stride(0), t.data_ptr(), t.stride(0), b, d, float(scale),
                                             mode, loss.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(),
                                             L.stream_ptr(p.device)))
    launch_counter.add(4 if mode == MODE_BF16 else 3)     # cast, GEMM+lse, combine, mean | GEMM, row lse, mean
    return loss, lse


def bbc_loss_backward(pred: torch.Tensor, tar: torch.Tensor, lse: torch.Tensor, grad_out: Optional[torch.Tensor] = None,
                      scale: float = 100.0, mode: int = MODE_BF16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gradient of :func:`bbc_loss_forward` with respect to ``pred`` and ``tar`` (both [B, D] fp32), multiplied by the
    device scalar ``grad_out`` (the upstream gradient, e.g. a GradScaler factor) when given."""
    p = _rowmajor(pred, "pred")
    t = _rowmajor(tar, "tar")
    L.require_cuda(lse, "lse")
    if p.dtype != torch.float32 or t.dtype != torch.float32 or p.shape != t.shape:
        raise ErnError("pred / tar must be float32 tensors of the same [B, D] shape")
    b, d = p.shape
    lse = lse.float().contiguous()
    if lse.numel() != b:
        raise ErnError("lse must hold one value per row")
    g_ptr = None
    if grad_out is not None:
        L.require_cuda(grad_out, "grad_out")
        grad_out = grad_out.reshape(-1).float().contiguous()
        g_ptr = grad_out.data_ptr()
    dpred = torch.empty((b, d), dtype=torch.float32, device=p.device)
    dtar = torch.empty((b, d), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        nbytes = L.lib().ern_bbc_loss_workspace_bytes(b, d, mode)
        ws = _workspace(nbytes, p.device)
        L.check(L.lib().ern_bbc_loss_backward(p.data_ptr(), p.stride(0), t.data_ptr(), t.stride(0), b, d, float(scale),
                                              mode, lse.data_ptr(), g_ptr, dpred.data_ptr(), dpred.stride(0),
                                              dtar.data_ptr(), dtar.stride(0), ws.data_ptr(), ws.numel(),
                                              L.stream_ptr(p.device)))
    launch_counter.add(5 if mode == MODE_BF16 else 6)     # cast+transpose, 2 softmax-gradient GEMMs, 2 GEMMs
    return dpred, dtar
