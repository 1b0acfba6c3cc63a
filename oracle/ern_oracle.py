"""CPU oracle for the composed-retrieval scoring path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 torch/numpy on the CPU, the algorithm of the reference
(ChenAnno/FashionERN_AAAI2024) for the one path this repository accelerates.  It is imported
only by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- never by the product package ``fashionern_aaai2024_b200`` (which has
no CPU fallback and raises when the CUDA library is missing).

Parity pin: ``oracle/make_golden.py`` runs the *unmodified* reference functions
(``/root/reference/run/test/test_{fiq,shoes,200k,cirr,val}.py`` and
``models/fusion_model.py``) on seeded synthetic inputs inside the authoring container, checks
that every function below returns bit-identical recall tuples / rankings on the same inputs, and
freezes the reference's outputs as fixtures under ``tests/golden/``.  ``tests/test_oracle.py``
re-checks this restatement against those fixtures on every run.

Reference lines followed by each function are cited in its docstring (paths relative to
``/root/reference``).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# Training criterion
# --------------------------------------------------------------------------------------
def bbc_loss(predicted_features: torch.Tensor, tar_features: torch.Tensor, scale: float = 100.0):
    """``BatchBasedClassificationLoss.forward`` (losses/loss.py:10-14) and the gradient autograd derives for it,
    written out explicitly in float64: logits = scale * P @ T.T (:11), labels = arange(B) (:12), mean cross entropy
    (:14);  dlogits = (softmax - I) / B, dP = scale * dlogits @ T, dT = scale * dlogits.T @ P.
    Returns (loss float, row logsumexp [B], dP [B,D], dT [B,D]) as float64 numpy."""
    p = predicted_features.detach().double().numpy()
    t = tar_features.detach().double().numpy()
    b = p.shape[0]
    x = scale * (p @ t.T)
    m = x.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(x - m).sum(axis=1, keepdims=True)))[:, 0]
    loss = float(np.mean(lse - np.diag(x)))
    dx = (np.exp(x - lse[:, None]) - np.eye(b)) / b
    return loss, lse, scale * (dx @ t), scale * (dx.T @ p)


# --------------------------------------------------------------------------------------
# Fusion head
# --------------------------------------------------------------------------------------
def combiner_forward(sd: Dict[str, torch.Tensor], image_features: torch.Tensor,
                     text_features: torch.Tensor, return_gate: bool = False):
    """``CombinerSimple.forward`` in eval mode -- models/fusion_model.py:86-94 (layers :73-84).

    Dropout (p=0.5, :76,:84) is the identity in eval.  Note the argument order (image, text) and
    that the concatenation puts the TEXT projection first (:90).
    """
    img = image_features.float()
    txt = text_features.float()
    tp = F.relu(F.linear(txt, sd["text_projection_layer.0.weight"], sd["text_projection_layer.0.bias"]))     # :87
    ip = F.relu(F.linear(img, sd["image_projection_layer.0.weight"], sd["image_projection_layer.0.bias"]))   # :88
    raw = torch.cat((tp, ip), dim=-1)                                                                         # :90
    h = F.relu(F.linear(raw, sd["dynamic_scalar.0.weight"], sd["dynamic_scalar.0.bias"]))                     # :74-75
    s = torch.sigmoid(F.linear(h, sd["dynamic_scalar.3.weight"], sd["dynamic_scalar.3.bias"]))                # :77-78
    out = s * txt + (1 - s) * img                                                                             # :93
    out = F.normalize(out, dim=-1)                                                                            # :94
    return (out, s) if return_gate else out


def visual_sr_forward(sd: Dict[str, torch.Tensor], local_feature: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``VisualSR.forward`` in eval mode -- models/fusion_model.py:141-154 (layers :106-124, l2norm :136-139).
    BatchNorm1d(13) sees [B,13,D]: channel = patch index; BatchNorm1d(D) sees [B,D]: channel = feature."""
    x = local_feature.float()
    raw_global = x.mean(dim=1)                                                                   # :142

    def bn(v, prefix, shape):
        m, var = sd[f"{prefix}.running_mean"].view(shape), sd[f"{prefix}.running_var"].view(shape)
        return (v - m) / torch.sqrt(var + eps) * sd[f"{prefix}.weight"].view(shape) + sd[f"{prefix}.bias"].view(shape)

    l_emb = torch.tanh(bn(F.linear(x, sd["embedding_local.0.weight"], sd["embedding_local.0.bias"]),
                          "embedding_local.1", (1, -1, 1)))                                      # :145
    g_emb = torch.tanh(bn(F.linear(raw_global, sd["embedding_global.0.weight"], sd["embedding_global.0.bias"]),
                          "embedding_global.1", (1, -1)))                                        # :146
    common = l_emb * g_emb.unsqueeze(1)                                                          # :149
    logits = F.linear(common, sd["embedding_common.weight"], sd["embedding_common.bias"]).squeeze(2)
    weights = torch.softmax(logits, dim=1)                                                       # :150
    new_global = (weights.unsqueeze(2) * x).sum(dim=1)                                           # :153
    norm = torch.sqrt((new_global ** 2).sum(dim=-1, keepdim=True)) + 1e-8                        # :136-139
    return new_global / norm


def _layer_norm(x, w, b, eps=1e-12):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def dvr_forward(sd: Dict[str, torch.Tensor], ref_patch_features: torch.Tensor, text_seq_features: torch.Tensor,
                ref_global_feats: torch.Tensor, text_global_feats: torch.Tensor, heads: int = 8,
                return_hidden: bool = False):
    """``DVR_module.forward`` in eval mode -- models/fusion_model.py:26-55, with ``PlusModel.forward`` (:187-216: HF
    BERT encoder over [CLS] + 13 patches + 77 tokens fed through ``inputs_embeds``) and the ``nn.MultiheadAttention``
    cross attention (:44-46) written out explicitly.  Dropouts are identities in eval."""
    B, P, D = ref_patch_features.shape
    T = text_seq_features.shape[1]
    L = 1 + P + T
    pre = "transformer_layer.bert_encoder.bert_model."
    x = torch.cat((sd["transformer_layer.cls_token"].expand(B, -1, -1), ref_patch_features.float(),
                   text_seq_features.float()), dim=1)                                              # :199-201
    tt = torch.cat((torch.zeros(P + 1, dtype=torch.long), torch.ones(T, dtype=torch.long)))        # :202-203
    x = x + sd[pre + "embeddings.token_type_embeddings.weight"][tt] + sd[pre + "embeddings.position_embeddings.weight"][:L]
    x = _layer_norm(x, sd[pre + "embeddings.LayerNorm.weight"], sd[pre + "embeddings.LayerNorm.bias"])
    dh = D // heads
    layer = 0
    while (pre + f"encoder.layer.{layer}.attention.self.query.weight") in sd:
        p = pre + f"encoder.layer.{layer}."
        q = F.linear(x, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"]).view(B, L, heads, dh).transpose(1, 2)
        k = F.linear(x, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"]).view(B, L, heads, dh).transpose(1, 2)
        v = F.linear(x, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"]).view(B, L, heads, dh).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2) / (dh ** 0.5), dim=-1)                          # attention_mask is all ones (:204)
        ctx = (a @ v).transpose(1, 2).reshape(B, L, D)
        x = _layer_norm(F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + x,
                        sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"])
        h = F.gelu(F.linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        x = _layer_norm(F.linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + x,
                        sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])
        layer += 1
    hidden = x                                                                                     # last_hidden_state (:213)
    image_norm = F.normalize(hidden[:, 1:P + 1], dim=2)                                            # :38-41
    text_norm = F.normalize(hidden[:, P + 1:], dim=2)
    w, bqkv = sd["MR_component.in_proj_weight"], sd["MR_component.in_proj_bias"]
    # only the first P query positions of the cross attention are used downstream (:47)
    q = F.linear(text_norm[:, :P], w[:D], bqkv[:D]).view(B, P, heads, dh).transpose(1, 2)
    k = F.linear(image_norm, w[D:2 * D], bqkv[D:2 * D]).view(B, P, heads, dh).transpose(1, 2)
    v = F.linear(image_norm, w[2 * D:], bqkv[2 * D:]).view(B, P, heads, dh).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) / (dh ** 0.5), dim=-1)
    cross = F.linear((a @ v).transpose(1, 2).reshape(B, P, D), sd["MR_component.out_proj.weight"], sd["MR_component.out_proj.bias"])
    sub = lambda prefix: {kk[len(prefix):]: vv for kk, vv in sd.items() if kk.startswith(prefix)}  # noqa: E731
    patch_vision_mean = visual_sr_forward(sub("SR_module."), cross)                               # :48
    seq_text_mean = text_norm.mean(dim=1)                                                          # :49
    global_feats = combiner_forward(sub("combiner_global."), ref_global_feats, text_global_feats)  # :52
    local_feats = combiner_forward(sub("combiner_local."), patch_vision_mean, seq_text_mean)      # :53
    fusion = combiner_forward(sub("combiner."), global_feats, local_feats)                         # :54
    return (fusion, hidden) if return_hidden else fusion


def gallery_normalize(index_features: torch.Tensor) -> torch.Tensor:
    """``F.normalize(index_features, dim=-1).float()`` -- run/test/test_fiq.py:45 (and twins)."""
    return F.normalize(index_features, dim=-1).float()


# --------------------------------------------------------------------------------------
# Scoring + ranking
# --------------------------------------------------------------------------------------
def distances(predicted_features: torch.Tensor, index_features: torch.Tensor) -> torch.Tensor:
    """``1 - predicted_features @ index_features.T`` -- run/test/test_fiq.py:49."""
    return 1 - predicted_features.float() @ index_features.float().T


def rank_topk(predicted_features: torch.Tensor, index_features: torch.Tensor, k: int,
              exclude_index: Optional[torch.Tensor] = None, chunk: int = 512
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """First ``k`` columns of ``torch.argsort(distances, dim=-1)`` -- run/test/test_fiq.py:49-50.

    The reference's argsort is unstable (tie order unspecified); here ties are broken by the lower
    gallery index, which is the rule the CUDA path implements.  ``exclude_index[q]`` (CIRR's
    reference image, run/test/test_cirr.py:55-58) is removed from row ``q`` before ranking.
    Returns (ids int64 [Q,k], dist fp32 [Q,k]); rows with fewer than k candidates are padded with
    id -1 / dist +inf.
    """
    q_total, n = predicted_features.shape[0], index_features.shape[0]
    kk = min(k, n)
    ids = torch.full((q_total, k), -1, dtype=torch.int64)
    dist = torch.full((q_total, k), float("inf"), dtype=torch.float32)
    g = index_features.float()
    for s in range(0, q_total, chunk):
        d = 1 - predicted_features[s:s + chunk].float() @ g.T
        if exclude_index is not None:
            ex = exclude_index[s:s + chunk].long()
            ok = ex >= 0
            rows = torch.arange(d.shape[0])[ok]
            d[rows, ex[ok]] = float("inf")
        sd, si = torch.sort(d, dim=-1, stable=True)
        ids[s:s + chunk, :kk] = si[:, :kk]
        dist[s:s + chunk, :kk] = sd[:, :kk]
    bad = torch.isinf(dist)
    ids[bad] = -1
    return ids, dist


# --------------------------------------------------------------------------------------
# Recall
# --------------------------------------------------------------------------------------
def percent(hits: int, total: int) -> float:
    """``(torch.sum(labels[:, :K]) / len(labels)).item() * 100`` -- run/test/test_fiq.py:59.

    int64 hit count and the Python int are both promoted to float32 before the division; the
    ``* 100`` happens in double after ``.item()``.
    """
    return float(np.float32(hits) / np.float32(total)) * 100


def factorize(index_names: Sequence[str], *others: Iterable[str]):
    """Map names to int class ids (same name -> same id).  Gallery position i gets class
    ``gallery_cls[i]``; names absent from the gallery map to -1."""
    table: Dict[str, int] = {}
    gallery_cls = np.empty(len(index_names), dtype=np.int64)
    for i, nm in enumerate(index_names):
        gallery_cls[i] = table.setdefault(nm, len(table))
    outs = [np.array([table.get(nm, -1) for nm in o], dtype=np.int64) for o in others]
    return (gallery_cls, *outs)


def first_hit_rank(top_ids: torch.Tensor, gallery_cls: np.ndarray, target_cls: np.ndarray) -> np.ndarray:
    """Rank (0-based) of the first ranked item whose class equals the target's; ``k`` if none.

    Unique names (run/test/test_fiq.py:51-55): exactly one gallery item matches, so this is the
    target's rank.  Fashion200k (run/test/test_200k.py:53-60): any-hit -- ``sum(labels[:, :K]) > 0``
    is ``first_hit_rank < K``.
    """
    ids = top_ids.numpy()
    k = ids.shape[1]
    cls = np.where(ids >= 0, gallery_cls[np.clip(ids, 0, None)], -2)
    hit = cls == target_cls[:, None]
    return np.where(hit.any(1), hit.argmax(1), k)


def recall_at(ranks: np.ndarray, ks: Sequence[int]) -> Tuple[float, ...]:
    q = len(ranks)
    return tuple(percent(int((ranks < k).sum()), q) for k in ks)


def check_unique_targets(index_names: Sequence[str], target_names: Sequence[str]) -> None:
    """The reference asserts every query's target name occurs exactly once in the ranked gallery
    (run/test/test_fiq.py:56); with names as given that is a statement about the name lists."""
    counts: Dict[str, int] = {}
    for nm in index_names:
        counts[nm] = counts.get(nm, 0) + 1
    for t in target_names:
        assert counts.get(t, 0) == 1


def fiq_metrics(pred, gallery, index_names, target_names, ks=(10, 50)) -> Tuple[float, ...]:
    """Tail of ``compute_fiq_val_metrics`` / ``compute_shoes_val_metrics`` -- run/test/test_fiq.py:48-64,
    run/test/test_shoes.py:47-61; with ``ks=(1,5,10,15,20,30,40,50)`` the VAL protocol of
    run/test/test_val.py:48-67."""
    check_unique_targets(index_names, target_names)
    gcls, tcls = factorize(index_names, target_names)
    ids, _ = rank_topk(pred, gallery, max(ks))
    return recall_at(first_hit_rank(ids, gcls, tcls), ks)


def f200k_metrics(pred, gallery, index_names, target_names, ks=(10, 50)) -> Tuple[float, ...]:
    """Tail of ``compute_200k_val_metrics`` -- run/test/test_200k.py:49-61 (non-unique caption names,
    any-hit recall; no cardinality assert)."""
    gcls, tcls = factorize(index_names, target_names)
    ids, _ = rank_topk(pred, gallery, max(ks))
    return recall_at(first_hit_rank(ids, gcls, tcls), ks)


def cirr_metrics(pred, gallery, index_names, reference_names, target_names, group_members
                 ) -> Tuple[float, ...]:
    """Tail of ``compute_cirr_val_metrics`` -- run/test/test_cirr.py:49-80.

    Reference removal (:55-58) == ranking with the reference's column excluded; subset recall
    (:64-66,:76-78) == rank of the target among the group members that survive the removal, ordered
    by the same distances.  Returns (G@1, G@2, G@3, R@1, R@5, R@10, R@50) (:80).
    """
    name_to_idx = {nm: i for i, nm in enumerate(index_names)}
    counts: Dict[str, int] = {}
    for nm in index_names:
        counts[nm] = counts.get(nm, 0) + 1
    q = len(target_names)
    for r, t in zip(reference_names, target_names):
        assert counts.get(r, 0) == 1 and counts.get(t, 0) == 1 and r != t          # :57-58,:68
    ref_idx = torch.tensor([name_to_idx[r] for r in reference_names])
    tgt_idx = np.array([name_to_idx[t] for t in target_names])
    ids, _ = rank_topk(pred, gallery, 50, exclude_index=ref_idx)
    gcls = np.arange(len(index_names))
    ranks = first_hit_rank(ids, gcls, tgt_idx)

    d = distances(pred, gallery)
    granks = np.empty(q, dtype=np.int64)
    for i in range(q):
        # group_mask marks ranked (reference-free) columns whose name is one of the 6 members (:64-65)
        mem = sorted({name_to_idx[m] for m in group_members[i] if m in name_to_idx and m != reference_names[i]})
        assert int(tgt_idx[i]) in mem                                                # :69
        dt = d[i, tgt_idx[i]]
        before = 0
        for m in mem:
            if m == tgt_idx[i]:
                continue
            dm = d[i, m]
            if dm < dt or (dm == dt and m < tgt_idx[i]):
                before += 1
        granks[i] = before
    g = recall_at(granks, (1, 2, 3))
    r = recall_at(ranks, (1, 5, 10, 50))
    return (*g, *r)


# --------------------------------------------------------------------------------------
# Tolerance-aware comparison of a candidate ranking with the oracle's
# --------------------------------------------------------------------------------------
def compare_topk(cand_ids: np.ndarray, cand_scores: Optional[np.ndarray], pred: torch.Tensor,
                 gallery: torch.Tensor, k: int, tol: float,
                 exclude_index: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """Check ``cand_ids`` ([Q,k] gallery indices, best first) against the oracle.

    Exact agreement is required except where the oracle's own scores are closer than ``tol``:
    at every rank j the item the candidate placed there must have an oracle similarity within
    ``tol`` of the item the oracle placed there, rows must be duplicate-free, and the reported
    score must match the oracle similarity of the reported item within ``tol``.
    Returns statistics; raises AssertionError on violation.
    """
    ids_o, dist_o = rank_topk(pred, gallery, k, exclude_index=exclude_index)
    sims = pred.float() @ gallery.float().T
    q = cand_ids.shape[0]
    cand = torch.from_numpy(np.ascontiguousarray(cand_ids)).long()
    valid = ids_o >= 0
    assert bool(((cand >= 0) == valid).all()), "padding pattern differs from oracle"
    srt, _ = torch.sort(torch.where(valid, cand, torch.arange(k).expand(q, k) - 10 - k), dim=1)
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "duplicate ids in a candidate row"
    s_c = torch.gather(sims, 1, cand.clamp(min=0))
    s_o = torch.gather(sims, 1, ids_o.clamp(min=0))
    gap = (s_c - s_o).abs()[valid]
    worst = float(gap.max()) if gap.numel() else 0.0
    assert worst <= tol, f"rank-wise oracle similarity differs by {worst} > tol {tol}"
    if exclude_index is not None:
        # padding (-1) never counts: a query without an exclusion carries exclude_index == -1
        assert not bool(((cand == exclude_index.long()[:, None]) & (cand >= 0)).any()), "excluded id was returned"
    out = {"exact_frac": float((cand == ids_o)[valid].float().mean()) if valid.any() else 1.0,
           "max_rank_gap": worst}
    if cand_scores is not None:
        err = (torch.from_numpy(np.ascontiguousarray(cand_scores)).float() - s_c).abs()[valid]
        out["max_score_err"] = float(err.max()) if err.numel() else 0.0
        assert out["max_score_err"] <= tol, f"reported score error {out['max_score_err']} > tol {tol}"
    return out
