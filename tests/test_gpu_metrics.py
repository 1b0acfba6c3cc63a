"""GPU: the drop-in compute_*_val_metrics against the recall tuples the UNMODIFIED reference produced."""
import numpy as np
import pytest
import torch

import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import metrics
from fashionern_aaai2024_b200.combiner import CombinerSimple
from helpers import FakeClip, FakeRelative, StandInERN, case_inputs, fake_tokenizer_factory, load_golden

pytestmark = pytest.mark.gpu

FN = {"fiq": metrics.compute_fiq_val_metrics, "val": metrics.compute_val_metrics,
      "shoes": metrics.compute_shoes_val_metrics, "200k": metrics.compute_200k_val_metrics,
      "cirr": metrics.compute_cirr_val_metrics}
KS = {"fiq": (10, 50), "shoes": (10, 50), "200k": (10, 50), "val": (1, 5, 10, 15, 20, 30, 40, 50),
      "cirr": (1, 2, 3, 1, 5, 10, 50)}


def build(name, dev, mode):
    z, meta = load_golden(name)
    index_features, index_local, names, states = case_inputs(meta)
    dim, kind = meta["dim"], meta["kind"]
    comb = CombinerSimple(dim, 4 * dim, 8 * dim, mode=mode)
    comb.load_state_dict(states["Combiner_module"])
    model = StandInERN(comb, torch.from_numpy(z["pred"]), torch.from_numpy(z["sr_out"])).to(dev).eval()
    ref_names = [names[i] for i in z["ref_idx"]]
    tgt_names = [names[i] for i in z["tgt_idx"]]
    members = [[names[m] for m in row] for row in z["members"]] if kind == "cirr" else None
    ds = FakeRelative("fiq" if kind == "val" else kind, ref_names, tgt_names, index_local[torch.from_numpy(z["ref_idx"]).long()], members)
    return z, meta, ds, index_features.to(dev), index_local.to(dev), names, model


@pytest.fixture(autouse=True)
def _tok():
    metrics.set_tokenizer_factory(fake_tokenizer_factory)
    yield
    metrics.set_tokenizer_factory(None)


@pytest.mark.parametrize("name", ["fiq640", "val512", "shoes640", "f200k640", "cirr640"])
def test_fp32_mode_reproduces_reference_tuple(cuda_device, name):
    z, meta, ds, feats, local, names, model = build(name, cuda_device, "fp32")
    out = FN[meta["kind"]](ds, FakeClip(meta["dim"]), feats, local, names, model, cuda_device, meta["dim"], 32, 0,
                           "RN50x4", precision="fp32")
    assert tuple(out) == tuple(z["recall"].tolist())     # bit-identical Recall@K percentages


@pytest.mark.parametrize("name", ["fiq640", "val512", "shoes640", "f200k640", "cirr640"])
def test_bf16_mode_is_exact_on_its_own_operands(cuda_device, name):
    """bf16 product path, tight: the drop-in function's tuple must EQUAL the oracle's tuple computed from the same
    bf16-rounded operands (the query features and the fused gallery the function itself scores), except for queries
    sitting inside a 2.2e-6 near-tie at a decision boundary (run/test/test_cirr.py:55-78 and twins).  The heads'
    bf16 error is bounded separately (tests/test_gpu_combiner.py); a loose check against the fp32 reference tuple
    guards against gross drift."""
    from helpers import assert_tuple_within_ties, cirr_oracle_with_ties, rounded, unique_oracle_with_ties
    z, meta, ds, feats, local, names, model = build(name, cuda_device, "bf16")
    kind, q, dim = meta["kind"], meta["q"], meta["dim"]
    out = FN[kind](ds, FakeClip(dim), feats, local, names, model, cuda_device, dim, 32, 0, "RN50x4", precision="bf16")
    pred_r = rounded(torch.from_numpy(z["pred"]))                      # StandInERN serves the recorded predictions
    gal_r = rounded(metrics.prepare_gallery(feats, local, model, cuda_device))
    if kind == "cirr":
        want, near = cirr_oracle_with_ties(pred_r, gal_r, names, ds.ref, ds.tgt, ds.members)
    else:
        want, near = unique_oracle_with_ties(pred_r, gal_r, names, ds.tgt, KS[kind], anyhit=(kind == "200k"))
    assert_tuple_within_ties(out, want, near, q)
    # vs the fp32 reference tuple: bf16 operand rounding (<= ~1e-3 per score) may move a query across a
    # K boundary only if its target sits that close to the boundary in the reference ranking
    d = z["ref_dist"]
    for j, (got, ref, k) in enumerate(zip(out, z["recall"].tolist(), KS[kind])):
        if kind == "cirr" and j < 3:
            continue                                                    # subset ranks: covered exactly above
        nr = int(np.sum(np.abs(d[:, min(k, d.shape[1] - 1)] - d[:, k - 1]) < 4e-3))
        assert abs(got - ref) <= 100.0 * nr / q + 1e-9, (k, got, ref, nr)


@pytest.mark.parametrize("name", ["fiq640", "f200k640", "cirr640"])
def test_fp16_tail_is_exact_on_its_own_operands(cuda_device, name):
    """precision="fp16": the same tensor-core kernels on fp16-rounded operands (ERN_DTYPE_F16).  The tuple must equal
    the oracle's on the same fp16-rounded operands except inside near-ties, as for bf16 above."""
    from helpers import assert_tuple_within_ties, cirr_oracle_with_ties, rounded, unique_oracle_with_ties
    z, meta, ds, feats, local, names, model = build(name, cuda_device, "bf16")
    kind, q, dim = meta["kind"], meta["q"], meta["dim"]
    out = FN[kind](ds, FakeClip(dim), feats, local, names, model, cuda_device, dim, 32, 0, "RN50x4", precision="fp16")
    pred_r = rounded(torch.from_numpy(z["pred"]), torch.float16)
    gal_r = rounded(metrics.prepare_gallery(feats, local, model, cuda_device), torch.float16)
    if kind == "cirr":
        want, near = cirr_oracle_with_ties(pred_r, gal_r, names, ds.ref, ds.tgt, ds.members)
    else:
        want, near = unique_oracle_with_ties(pred_r, gal_r, names, ds.tgt, KS[kind], anyhit=(kind == "200k"))
    assert_tuple_within_ties(out, want, near, q)


def test_8_argument_form_and_print(cuda_device, capsys):
    z, meta, ds, feats, local, names, model = build("fiq640", cuda_device, "fp32")
    metrics.set_precision("fp32")
    try:
        out = metrics.compute_fiq_val_metrics(ds, FakeClip(640), feats, local, names, model, cuda_device, 640)
    finally:
        metrics.set_precision("bf16")
    assert tuple(out) == tuple(z["recall"].tolist())
    assert "R@10:" in capsys.readouterr().out            # run/test/test_fiq.py:62


def test_error_conventions(cuda_device):
    z, meta, ds, feats, local, names, model = build("fiq640", cuda_device, "fp32")
    bad = list(names)
    bad[int(z["tgt_idx"][0])] = "not-in-gallery"          # target 0 no longer matches any gallery name
    ds.ref = [n if n in bad else bad[1] for n in ds.ref]
    with pytest.raises(AssertionError):
        metrics.compute_fiq_val_metrics(ds, FakeClip(640), feats, local, bad, model, cuda_device, 640, precision="fp32")
    z, meta, ds, feats, local, names, model = build("cirr640", cuda_device, "fp32")
    ds.members = [[m for m in row if m != t] + [row[0]] for row, t in zip(ds.members, ds.tgt)]   # target not in group
    with pytest.raises(AssertionError):
        metrics.compute_cirr_val_metrics(ds, FakeClip(640), feats, local, names, model, cuda_device, 640, precision="fp32")


def test_full_dataset_shape_against_oracle(cuda_device):
    # FashionIQ-dress shape (2017 x 3817 x 640): tail only, synthetic unit features, planted targets
    from oracle import ern_oracle as orc
    from fashionern_aaai2024_b200 import synthetic as syn
    q, n, dim = 2017, 3817, 640
    pred, gal = syn.features(41, q, dim, unit=True), syn.features(42, n, dim, unit=True)
    ids_o, _ = orc.rank_topk(pred, gal, 51)
    tgt = ids_o[torch.arange(q), syn.planted_ranks(43, q, 51).clamp(max=50)]
    names = syn.unique_names(n)
    want = orc.fiq_metrics(pred, gal, names, [names[int(i)] for i in tgt], (10, 50))
    got = ern.score_topk_recall(pred.to(cuda_device), gal.to(cuda_device),
                                torch.arange(n, dtype=torch.int32, device=cuda_device),
                                tgt.int().to(cuda_device), (10, 50), precision="fp32")
    assert got["recall"] == want
