"""Drop-in replacements of the reference's validation-metric functions (the scoring tail).

Mirrors, with the same names, positional arguments, return tuples and error behaviour:

  compute_fiq_val_metrics    run/test/test_fiq.py:18-64     (8-arg twin run/valid/validate_fiq.py:11-47)
  compute_shoes_val_metrics  run/test/test_shoes.py:18-61   (run/valid/validate_shoes.py:11-48)
  compute_200k_val_metrics   run/test/test_200k.py:20-61
  compute_cirr_val_metrics   run/test/test_cirr.py:18-80    (run/valid/validate_cirr.py:11-72)
  compute_val_metrics        run/test/test_val.py:18-67     (the 8-K "VAL" protocol; named
                             compute_fiq_val_metrics in that file)
  generate_*_val_predictions run/test/test_fiq.py:67-122 and twins (the caller that feeds the tail)

What changes underneath: the tail no longer materialises ``1 - pred @ index.T`` nor fully argsorts it,
no [Q,N] index matrix crosses to the host and no numpy string gather/compare happens.  Instead

  gallery L2-normalise ........ ern_l2norm_rows          (test_fiq.py:45)
  model(mode="index") ......... the caller's ERN; its combiners are the B200 head after
                                ``accelerate_ern(model)`` (test_fiq.py:46 -> models/model.py:64-66)
  similarity + top-k .......... ern_sim_topk             (test_fiq.py:49-50), K_max columns only
  Recall@K / any-hit .......... ern_recall_at_k          (test_fiq.py:51-60, test_200k.py:52-60)
  CIRR reference removal ...... exclude ids in ern_sim_topk (test_cirr.py:55-58)
  CIRR subset recall .......... ern_cirr_subset_recall   (test_cirr.py:64-66,76-78)

Names (Python strings) are factorised once on the host into int32 ids; only hit counts (a few ints) come
back from the device.  ``precision`` selects the arithmetic: "bf16" (tensor cores, default), "fp16" (the same
tensor-core kernel on fp16-rounded operands: 11 mantissa bits instead of 8 at the same rate, safe for the unit-norm
features this tail scores) or "fp32" (validation mode that reproduces the reference's fp32 ranking of ``1 - s``
including its rounding ties).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import ops
from ._lib import MODE_BF16, MODE_FP32, RANK_REFERENCE, ErnError, require_cuda

_PRECISION = "bf16"
_PRECISIONS = ("bf16", "fp16", "fp32")
_TOKENIZER_FACTORY: Optional[Callable[[str], Callable]] = None


def set_precision(precision: str) -> None:
    """Default arithmetic of the tail: 'bf16' (product path), 'fp16' (same kernel, fp16 operands) or 'fp32'
    (validation mode)."""
    global _PRECISION
    if precision not in _PRECISIONS:
        raise ErnError(f"unknown precision {precision!r}")
    _PRECISION = precision


def set_tokenizer_factory(factory: Optional[Callable[[str], Callable]]) -> None:
    """Override ``open_clip.get_tokenizer`` (run/test/test_fiq.py:79), e.g. when open_clip is not installed."""
    global _TOKENIZER_FACTORY
    _TOKENIZER_FACTORY = factory


def _tokenizer(clip_model_name: str):
    if _TOKENIZER_FACTORY is not None:
        return _TOKENIZER_FACTORY(clip_model_name)
    try:
        import open_clip  # noqa: WPS433 (the reference's own dependency)
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ErnError("open_clip is not installed; install it or call set_tokenizer_factory()") from e
    return open_clip.get_tokenizer(clip_model_name)


def collate_fn(batch: list):
    """Drop ``None`` samples, then default-collate (utils/utils.py:22-29)."""
    batch = [b for b in batch if b is not None]
    return torch.utils.data.dataloader.default_collate(batch)


def percent(hits: int, total: int) -> float:
    """``(torch.sum(labels[:, :K]) / len(labels)).item() * 100`` (run/test/test_fiq.py:59): the division is
    carried out in float32, the scaling in double."""
    return float(np.float32(hits) / np.float32(total)) * 100


# ---------------------------------------------------------------------------------------------------
# Query-side feature production (the caller of the hot path; host logic mirroring the reference)
# ---------------------------------------------------------------------------------------------------
def _generate(kind: str, clip_model, relative_val_dataset, model, index_names, index_features, device,
              feature_dim, batch_size, num_workers, clip_model_name):
    tokenizer = _tokenizer(clip_model_name)
    loader = DataLoader(dataset=relative_val_dataset, batch_size=batch_size, num_workers=num_workers,
                        pin_memory=True, collate_fn=collate_fn, shuffle=False)
    # dict(zip(names, rows)) keeps the LAST row of a repeated name (run/test/test_fiq.py:88; matters for
    # Fashion200k's caption names); rows are gathered on the device by index instead of itemgetter + stack
    name_to_row: Dict[str, int] = {nm: i for i, nm in enumerate(index_names)}
    index_features = index_features.to(device)
    preds: List[torch.Tensor] = []
    target_names: List[str] = []
    reference_names: List[str] = []
    group_members: List[List[str]] = []
    for batch in loader:
        if kind == "fiq":
            ref_names, batch_targets, captions, ref_patch = batch
            texts = [f"{a.strip('.?, ').capitalize()} and {b.strip('.?, ')}" for a, b in zip(captions[0], captions[1])]
        elif kind == "shoes":
            ref_names, batch_targets, texts, ref_patch, _ = batch
        elif kind == "cirr":
            ref_names, batch_targets, texts, ref_patch, members = batch
            group_members.extend(np.array(members).T.tolist())
        elif kind == "200k":
            _, ref_names, texts, batch_targets, _, ref_patch = batch
        else:
            raise ErnError(f"unknown dataset kind {kind}")
        text_inputs = tokenizer(list(texts), context_length=77).to(device)
        ref_patch = ref_patch.to(device)
        with torch.no_grad():
            visual = ref_patch.transpose(0, 1)
            text_features, _ = clip_model.encode_text(text_inputs, visual_emb=visual)
            text_seq = clip_model.encode_text(text_inputs, mode="seq", visual_emb=visual)
            rows = torch.tensor([name_to_row[n] for n in ref_names], device=device, dtype=torch.long)
            ref_feats = index_features.index_select(0, rows)
            out = model(ref_feats=ref_feats, ref_local_feats=ref_patch, text_feats=text_features.to(device),
                        text_seq_feats=text_seq.to(device), mode="test")
        preds.append(out)
        target_names.extend(batch_targets)
        reference_names.extend(ref_names)
    pred = torch.cat(preds, 0) if preds else torch.empty((0, feature_dim), device=device)
    return pred, reference_names, target_names, group_members


def generate_fiq_val_predictions(clip_model, relative_val_dataset, model, index_names, index_features, device,
                                 feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4"):
    """run/test/test_fiq.py:67-122 -> (predicted_features [Q,D], target_names)."""
    p, _, t, _ = _generate("fiq", clip_model, relative_val_dataset, model, index_names, index_features, device,
                           feature_dim, batch_size, num_workers, clip_model_name)
    return p, t


def generate_shoes_val_predictions(clip_model, relative_val_dataset, model, index_names, index_features, device,
                                   feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4"):
    """run/test/test_shoes.py:64-118 -> (predicted_features, target_names)."""
    p, _, t, _ = _generate("shoes", clip_model, relative_val_dataset, model, index_names, index_features, device,
                           feature_dim, batch_size, num_workers, clip_model_name)
    return p, t


def generate_200k_val_predictions(clip_model, relative_val_dataset, model, index_names, index_features, device,
                                  feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4"):
    """run/test/test_200k.py:64-113 -> (predicted_features, target_names); kept on the device here (the
    reference moves them to the CPU and scores there, :48,:86,:111)."""
    p, _, t, _ = _generate("200k", clip_model, relative_val_dataset, model, index_names, index_features, device,
                           feature_dim, batch_size, num_workers, clip_model_name)
    return p, t


def generate_cirr_val_predictions(clip_model, relative_val_dataset, model, index_names, index_features, device,
                                  feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4"):
    """run/test/test_cirr.py:83-139 -> (predicted_features, reference_names, target_names, group_members)."""
    return _generate("cirr", clip_model, relative_val_dataset, model, index_names, index_features, device,
                     feature_dim, batch_size, num_workers, clip_model_name)


# ---------------------------------------------------------------------------------------------------
# The tail on the device
# ---------------------------------------------------------------------------------------------------
def factorize_names(index_names: Sequence[str]):
    """names -> (class id per gallery row int32 [N], table name->class id, occurrences per class)."""
    table: Dict[str, int] = {}
    cls = np.empty(len(index_names), dtype=np.int32)
    for i, nm in enumerate(index_names):
        cls[i] = table.setdefault(nm, len(table))
    counts = np.bincount(cls, minlength=len(table)) if len(cls) else np.zeros(0, dtype=np.int64)
    return cls, table, counts


def prepare_gallery(index_features: torch.Tensor, index_local_features, model, device) -> torch.Tensor:
    """``F.normalize(index_features).float()`` then ``model(mode="index").float()``
    (run/test/test_fiq.py:45-46 -> models/model.py:64-66)."""
    index_features = index_features.to(device)
    normed, _ = ops.l2norm_rows(index_features.float(), normalize=True)
    with torch.no_grad():
        gallery = model(tar_feats=normed, tar_local_feats=index_local_features, mode="index").float()
    return gallery


def _operands(pred: torch.Tensor, gallery: torch.Tensor, precision: str):
    pred = pred.float().contiguous()
    gallery = gallery.float().contiguous()
    if precision not in _PRECISIONS:
        raise ErnError(f"unknown precision {precision!r}")
    if precision == "fp32":
        return pred, gallery, MODE_FP32
    # 16-bit operands for the tensor-core path; pad D up to a multiple of 64 with zeros (cosine unchanged)
    d = pred.shape[1]
    dp = (d + 63) // 64 * 64
    if dp != d:
        pred = torch.nn.functional.pad(pred, (0, dp - d))
        gallery = torch.nn.functional.pad(gallery, (0, dp - d))
    f16 = precision == "fp16"
    _, qb = ops.l2norm_rows(pred, normalize=False, want_f32=False, want_bf16=not f16, want_f16=f16)
    _, gb = ops.l2norm_rows(gallery, normalize=False, want_f32=False, want_bf16=not f16, want_f16=f16)
    return qb, gb, MODE_BF16


def score_topk_recall(predicted_features: torch.Tensor, gallery_features: torch.Tensor, gallery_class: torch.Tensor,
                      target_class: torch.Tensor, ks: Sequence[int], precision: Optional[str] = None,
                      exclude_ids: Optional[torch.Tensor] = None, k: Optional[int] = None) -> Dict[str, object]:
    """The tail on its own, for callers without a Dataset (synthetic scaling config):
    cosine top-k of every query against the gallery + Recall@K from id membership.

    ``gallery_class`` int32 [N] (class/name id per gallery row), ``target_class`` int32 [Q].  Returns
    ``{"hits": [..], "recall": (percent per K), "ranks": int32[Q] (device), "top_ids", "top_values"}``.
    """
    precision = precision or _PRECISION
    require_cuda(predicted_features, "predicted_features")
    require_cuda(gallery_features, "gallery_features")
    q, g, mode = _operands(predicted_features, gallery_features, precision)
    kk = int(k or max(ks))
    vals, ids, _, _ = ops.sim_topk(q, g, kk, mode=mode, rank_by=RANK_REFERENCE, exclude_ids=exclude_ids)
    counts, ranks = ops.recall_at_k(ids, gallery_class.to(torch.int32), target_class.to(torch.int32), ks)
    hits = counts.cpu().tolist()
    nq = predicted_features.shape[0]
    return {"hits": hits, "recall": tuple(percent(h, nq) for h in hits), "ranks": ranks, "top_ids": ids,
            "top_values": vals}


def _unique_tail(pred, gallery, index_names, target_names, ks, precision, device):
    cls, table, counts = factorize_names(index_names)
    tcls = np.array([table.get(t, -1) for t in target_names], dtype=np.int32)
    # "every target name occurs exactly once in the ranked gallery" (run/test/test_fiq.py:56)
    assert bool(np.all(tcls >= 0)) and bool(np.all(counts[tcls] == 1))
    res = score_topk_recall(pred, gallery, torch.from_numpy(cls).to(device), torch.from_numpy(tcls).to(device), ks,
                            precision)
    return res["recall"]


def compute_fiq_val_metrics(relative_val_dataset, clip_model, index_features, index_local_features, index_names,
                            model, device, feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4",
                            *, precision: Optional[str] = None) -> Tuple[float, float]:
    """(Recall@10, Recall@50) -- run/test/test_fiq.py:18-64."""
    pred, target_names = generate_fiq_val_predictions(clip_model, relative_val_dataset, model, index_names,
                                                      index_features, device, feature_dim, batch_size, num_workers,
                                                      clip_model_name)
    gallery = prepare_gallery(index_features, index_local_features, model, device)
    r10, r50 = _unique_tail(pred, gallery, index_names, target_names, (10, 50), precision, device)
    print("R@10:", r10, "   R@50:", r50)  # run/test/test_fiq.py:62
    return r10, r50


def compute_shoes_val_metrics(relative_val_dataset, clip_model, index_features, index_local_features, index_names,
                              model, device, feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4",
                              *, precision: Optional[str] = None) -> Tuple[float, float]:
    """(Recall@10, Recall@50) -- run/test/test_shoes.py:18-61."""
    pred, target_names = generate_shoes_val_predictions(clip_model, relative_val_dataset, model, index_names,
                                                        index_features, device, feature_dim, batch_size, num_workers,
                                                        clip_model_name)
    gallery = prepare_gallery(index_features, index_local_features, model, device)
    return _unique_tail(pred, gallery, index_names, target_names, (10, 50), precision, device)


def compute_val_metrics(relative_val_dataset, clip_model, index_features, index_local_features, index_names,
                        model, device, feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4",
                        *, precision: Optional[str] = None) -> Tuple[float, ...]:
    """Recall@{1,5,10,15,20,30,40,50} -- the VAL protocol of run/test/test_val.py:18-67."""
    pred, target_names = generate_fiq_val_predictions(clip_model, relative_val_dataset, model, index_names,
                                                      index_features, device, feature_dim, batch_size, num_workers,
                                                      clip_model_name)
    gallery = prepare_gallery(index_features, index_local_features, model, device)
    return _unique_tail(pred, gallery, index_names, target_names, (1, 5, 10, 15, 20, 30, 40, 50), precision, device)


def compute_200k_val_metrics(relative_val_dataset, clip_model, index_features, index_local_features, index_names,
                             model, device, feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4",
                             *, precision: Optional[str] = None, ks: Sequence[int] = (10, 50)) -> Tuple[float, ...]:
    """(Recall@10, Recall@50), any-hit over non-unique caption names -- run/test/test_200k.py:20-61.
    ``ks=(1, 10, 50)`` adds the R@1 that BASELINE.json's config asks for (same any-hit rule)."""
    pred, target_names = generate_200k_val_predictions(clip_model, relative_val_dataset, model, index_names,
                                                       index_features, device, feature_dim, batch_size, num_workers,
                                                       clip_model_name)
    gallery = prepare_gallery(index_features, index_local_features, model, device)
    cls, table, _ = factorize_names(index_names)
    tcls = np.array([table.get(t, -1) for t in target_names], dtype=np.int32)
    res = score_topk_recall(pred, gallery, torch.from_numpy(cls).to(device), torch.from_numpy(tcls).to(device), ks,
                            precision)
    return res["recall"]


def cirr_tail(pred, gallery, index_names, reference_names, target_names, group_members, precision, device):
    """run/test/test_cirr.py:49-80 on the device -> (G@1, G@2, G@3, R@1, R@5, R@10, R@50)."""
    precision = precision or _PRECISION
    cls, table, counts = factorize_names(index_names)
    nq = len(target_names)
    ref_cls = np.array([table.get(r, -1) for r in reference_names], dtype=np.int64)
    tgt_cls = np.array([table.get(t, -1) for t in target_names], dtype=np.int64)
    if not (np.all(ref_cls >= 0) and np.all(counts[ref_cls] == 1)):
        # the reference's reshape to [Q, N-1] fails when a reference image is not matched exactly once (:57-58)
        raise ValueError("cannot remove the reference image: reference name not matched exactly once in index_names")
    assert bool(np.all(tgt_cls >= 0)) and bool(np.all(counts[tgt_cls] == 1)) and bool(np.all(tgt_cls != ref_cls))  # :68
    first_row = np.full(len(table), -1, dtype=np.int64)
    first_row[cls[::-1]] = np.arange(len(cls))[::-1]
    ref_row = first_row[ref_cls].astype(np.int32)
    tgt_row = first_row[tgt_cls].astype(np.int32)
    mem = np.array([[first_row[table[m]] if m in table else -1 for m in row] for row in group_members], dtype=np.int32)
    mem = mem.reshape(nq, -1)

    q, g, mode = _operands(pred, gallery, precision)
    ref_dev = torch.from_numpy(ref_row).to(device)
    tgt_dev = torch.from_numpy(tgt_row).to(device)
    vals, ids, _, _ = ops.sim_topk(q, g, 50, mode=mode, rank_by=RANK_REFERENCE, exclude_ids=ref_dev)
    rcounts, _ = ops.recall_at_k(ids, torch.arange(len(cls), dtype=torch.int32, device=device), tgt_dev, (1, 5, 10, 50))
    gcounts, granks = ops.cirr_subset_recall(q, g, torch.from_numpy(mem).to(device), ref_dev, tgt_dev, (1, 2, 3),
                                             rank_by=RANK_REFERENCE)
    assert bool((granks >= 0).all().item())  # target must be one of the surviving group members (:69)
    r = [percent(h, nq) for h in rcounts.cpu().tolist()]
    gr = [percent(h, nq) for h in gcounts.cpu().tolist()]
    return (gr[0], gr[1], gr[2], r[0], r[1], r[2], r[3])


def compute_cirr_val_metrics(relative_val_dataset, clip_model, index_features, index_local_features, index_names,
                             model, device, feature_dim, batch_size=32, num_workers=4, clip_model_name="RN50x4",
                             *, precision: Optional[str] = None) -> Tuple[float, ...]:
    """(G@1, G@2, G@3, R@1, R@5, R@10, R@50) -- run/test/test_cirr.py:18-80."""
    pred, reference_names, target_names, group_members = generate_cirr_val_predictions(
        clip_model, relative_val_dataset, model, index_names, index_features, device, feature_dim, batch_size,
        num_workers, clip_model_name)
    gallery = prepare_gallery(index_features, index_local_features, model, device)
    return cirr_tail(pred, gallery, index_names, reference_names, target_names, group_members, precision, device)
