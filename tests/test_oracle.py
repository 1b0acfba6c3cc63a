"""CPU: the oracle restatement against the fixtures frozen from the UNMODIFIED reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ern_oracle as orc
from fashionern_aaai2024_b200 import synthetic as syn
from helpers import GOLDEN, case_inputs, load_golden

CASES = ["fiq640", "val512", "shoes640", "f200k640", "cirr640"]


@pytest.mark.parametrize("dim", [640, 512])
def test_combiner_restatement_matches_reference(dim):
    z, meta = load_golden(f"combiner{dim}")
    sd = syn.combiner_state(meta["seed"], dim)
    for case in ("raw", "unit", "zero"):
        out = orc.combiner_forward(sd, torch.from_numpy(z[f"{case}_image"]), torch.from_numpy(z[f"{case}_text"]))
        ref = torch.from_numpy(z[f"{case}_out"])
        assert torch.allclose(out, ref, atol=1e-6, rtol=0), case
        n = out.norm(dim=-1)
        if case != "zero":
            assert torch.allclose(n, torch.ones_like(n), atol=1e-5)
        else:
            assert torch.equal(out, torch.zeros_like(out))  # F.normalize of a zero row is zero (eps clamp)


@pytest.mark.parametrize("name", CASES)
def test_ranking_and_recall_match_reference(name):
    z, meta = load_golden(name)
    _, _, names, _ = case_inputs(meta)
    pred, gallery = torch.from_numpy(z["pred"]), torch.from_numpy(z["gallery"])
    ref_top = torch.from_numpy(z["ref_top"]).long()
    ids, dist = orc.rank_topk(pred, gallery, ref_top.shape[1])
    # same distances rank by rank; ids may only differ inside exact ties of the reference's fp32 distances
    assert np.array_equal(dist.numpy(), z["ref_dist"])
    assert int((ids != ref_top).sum()) == meta["tie_positions"]
    tgt_names = [names[i] for i in z["tgt_idx"]]
    kind = meta["kind"]
    if kind in ("fiq", "shoes"):
        mine = orc.fiq_metrics(pred, gallery, names, tgt_names, (10, 50))
    elif kind == "val":
        mine = orc.fiq_metrics(pred, gallery, names, tgt_names, (1, 5, 10, 15, 20, 30, 40, 50))
    elif kind == "200k":
        mine = orc.f200k_metrics(pred, gallery, names, tgt_names, (10, 50))
    else:
        ref_names = [names[i] for i in z["ref_idx"]]
        members = [[names[m] for m in row] for row in z["members"]]
        mine = orc.cirr_metrics(pred, gallery, names, ref_names, tgt_names, members)
    assert tuple(mine) == tuple(z["recall"].tolist())          # bit-identical percentages
    assert any(0 < r < 100 for r in mine)                      # planted targets make the metric non-trivial


def test_full_size_pins_recorded():
    with open(os.path.join(GOLDEN, "pin_report.json")) as f:
        rep = json.load(f)
    # make_golden.py asserts restatement == reference before writing these entries
    assert rep["full"]["fiq_dress_full"]["q"] == 2017 and rep["full"]["fiq_dress_full"]["n"] == 3817
    assert rep["full"]["cirr_val_full"]["q"] == 4181 and rep["full"]["cirr_val_full"]["n"] == 2297
    assert len(rep["full"]["cirr_val_full"]["recall"]) == 7


def test_percent_is_float32_division():
    # (torch.sum(labels[:, :K]) / len(labels)).item() * 100 -- SURVEY.md R7
    assert orc.percent(7, 2017) == 0.3470500698313117
    t = (torch.tensor(7) / 2017).item() * 100
    assert orc.percent(7, 2017) == t


def test_assert_conventions():
    names = ["a", "b", "b", "c"]
    g = torch.eye(4)
    p = torch.eye(4)[:2]
    with pytest.raises(AssertionError):
        orc.fiq_metrics(p, g, names, ["a", "b"])        # duplicated target name
    with pytest.raises(AssertionError):
        orc.fiq_metrics(p, g, names, ["a", "zzz"])      # missing target
    assert orc.f200k_metrics(p, g, names, ["a", "b"], (1,)) == (100.0,)   # any-hit tolerates duplicates


def test_compare_topk_tolerance():
    g = torch.nn.functional.normalize(torch.randn(50, 16, generator=torch.Generator().manual_seed(0)), dim=-1)
    p = g[:5] + 0.01
    ids, _ = orc.rank_topk(p, g, 10)
    stats = orc.compare_topk(ids.numpy(), None, p, g, 10, tol=0.0)
    assert stats["exact_frac"] == 1.0
    bad = ids.numpy().copy()
    bad[0, 0], bad[0, 9] = bad[0, 9], bad[0, 0]
    with pytest.raises(AssertionError):
        orc.compare_topk(bad, None, p, g, 10, tol=1e-6)


@pytest.mark.parametrize("dim", [640, 512])
def test_visualsr_restatement_matches_reference(dim):
    z, meta = load_golden(f"visualsr{dim}")
    sd = syn.visualsr_state(meta["seed"], dim)
    x = syn.patch_features(meta["seed"] + 1, meta["rows"], dim)
    out = orc.visual_sr_forward(sd, x)
    assert torch.allclose(out, torch.from_numpy(z["out"]), atol=2e-7, rtol=0)
    n = out.norm(dim=-1)
    assert torch.allclose(n, torch.ones_like(n), atol=1e-5)


@pytest.mark.parametrize("dim", [640, 512])
def test_dvr_restatement_matches_reference(dim):
    # reference DVR_module (HF BERT + nn.MultiheadAttention + VisualSR + 3 heads) vs the explicit restatement
    z, meta = load_golden(f"dvr{dim}")
    sd = syn.dvr_full_state(meta["seed"], dim)
    s, rows = meta["seed"], meta["rows"]
    out, hidden = orc.dvr_forward(sd, syn.patch_features(s + 10, rows, dim), syn.token_features(s + 11, rows, dim),
                                  syn.features(s + 12, rows, dim), syn.features(s + 13, rows, dim), return_hidden=True)
    assert float((out - torch.from_numpy(z["out"])).abs().max()) < 2e-6
    assert float((hidden[:, 0] - torch.from_numpy(z["hidden_cls"])).abs().max()) < 2e-5
    assert float((hidden[:, -1] - torch.from_numpy(z["hidden_last"])).abs().max()) < 2e-5


@pytest.mark.parametrize("dim", [640, 512])
def test_loss_restatement_matches_reference(dim):
    # reference BatchBasedClassificationLoss + torch autograd (losses/loss.py:10-14) vs the explicit float64 formulas
    z, meta = load_golden(f"bbcloss{dim}")
    pred, tar = syn.loss_pair(meta["seed"], meta["rows"], dim)
    loss, lse, dp, dt = orc.bbc_loss(pred, tar)
    assert abs(loss - float(z["loss"])) <= 5e-6 * abs(float(z["loss"]))
    assert 0.05 < loss < 2.0                                  # neither saturated nor at chance (log B)
    assert float(np.abs(dp - z["dpred"]).max()) < 5e-6 and float(np.abs(dt - z["dtar"]).max()) < 5e-6
    # gradient of a mean of cross entropies: every column of dlogits sums to a known value -> sum_i dT rows
    eps = 1e-3
    d = torch.zeros_like(pred)
    d[3, 5] = eps
    num = (orc.bbc_loss(pred + d, tar)[0] - orc.bbc_loss(pred - d, tar)[0]) / (2 * eps)
    assert abs(num - dp[3, 5]) < 1e-6 + 1e-4 * abs(dp[3, 5])


def test_compare_topk_padding_is_not_an_excluded_id():
    # gallery smaller than k: rows are padded with -1, and a query WITHOUT an exclusion carries exclude_index == -1
    p, g = syn.features(1, 3, 16, unit=True), syn.features(2, 2, 16, unit=True)
    excl = torch.tensor([-1, 0, 1])
    ids, _ = orc.rank_topk(p, g, 5, exclude_index=excl)
    assert (ids[0] >= 0).sum() == 2 and (ids[1] >= 0).sum() == 1
    orc.compare_topk(ids.numpy(), None, p, g, 5, tol=1e-6, exclude_index=excl)
    bad = ids.clone()
    bad[1, 1] = 0                                   # returns the excluded row 0 for query 1
    with pytest.raises(AssertionError):
        orc.compare_topk(bad.numpy(), None, p, g, 5, tol=1e-6, exclude_index=excl)
