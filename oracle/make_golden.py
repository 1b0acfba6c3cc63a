"""Mint the parity fixtures in ``tests/golden/`` by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run inside the authoring container (needs ``/root/reference``):

    python oracle/make_golden.py            # small committed fixtures + full-size pin report
    python oracle/make_golden.py --no-full  # skip the full-size (2017x3817, 4181x2297) pin runs

For every dataset kind it
  1. builds seeded synthetic inputs (``fashionern_aaai2024_b200.synthetic``),
  2. calls the reference's ``compute_*_val_metrics`` verbatim (``oracle/ref_harness.py``), recording the
     tensors that cross the hot-path boundary (query features, VisualSR output, fused gallery, the
     reference's own ``torch.argsort`` result),
  3. re-plants each query's target at a chosen rank of the reference's ranking (forced mass at the K
     boundaries) and calls the reference again -> the golden recall tuple,
  4. asserts that the CPU restatement in ``oracle/ern_oracle.py`` reproduces the reference's recall tuple
     bit for bit and its ranking everywhere outside exact distance ties,
  5. writes ``tests/golden/<case>.npz``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fashionern_aaai2024_b200 import synthetic as syn  # noqa: E402
from oracle import ern_oracle as orc  # noqa: E402
from oracle import ref_harness as ref  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TOPC = 51  # columns of the reference ranking that are frozen (K_max + 1)

COMBINERS = ("DVR.combiner_global", "DVR.combiner_local", "DVR.combiner", "Combiner_module")


def case_inputs(kind: str, dim: int, q: int, n: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    d = {
        "index_features": syn.features(seed + 1, n, dim),
        "index_local": syn.patch_features(seed + 2, n, dim),
        "text_global": syn.features(seed + 3, q, dim),
        "text_seq": syn.token_features(seed + 4, q, dim),
        "ref_idx": torch.randint(0, n, (q,), generator=g),
        "rand_tgt": torch.randint(0, n, (q,), generator=g),
    }
    if kind == "200k":
        d["names"] = syn.caption_names(seed + 6, n, classes=max(8, n // 6))
    elif kind == "cirr":
        d["names"] = syn.unique_names(n, "dev-{}-img")
    else:
        d["names"] = syn.unique_names(n)
    d["states"] = {name: syn.combiner_state(seed + 10 + i, dim) for i, name in enumerate(COMBINERS)}
    return d


def cirr_members(seed, q, n, ref_idx, tgt_idx):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(q):
        r, t = int(ref_idx[i]), int(tgt_idx[i])
        others = [x for x in torch.randperm(n, generator=g).tolist() if x != r and x != t][:4]
        mem = [r, t] + others
        perm = torch.randperm(6, generator=g).tolist()
        out.append([mem[p] for p in perm])
    return out


def run_reference(kind, dim, inp, ref_idx, tgt_idx, members_idx, model, clip):
    names = inp["names"]
    ref_names = [names[int(i)] for i in ref_idx]
    tgt_names = [names[int(i)] for i in tgt_idx]
    members = None if members_idx is None else [[names[m] for m in row] for row in members_idx]
    # each ref patch is the reference image's local feature
    ds = ref.FakeRelative("fiq" if kind == "val" else kind, ref_names, tgt_names,
                          inp["index_local"][ref_idx], members)
    rec = ref.Recorder(model)
    store = []
    with ref.capture_argsort(store):
        res = ref.run_metric(kind, ds, clip, inp["index_features"], inp["index_local"], names, rec, dim)
    pred = torch.cat(rec.pred, 0)
    return res, pred, rec, store[0], ref_names, tgt_names, members


def make_case(name, kind, dim, q, n, seed, save=True, hooks=True):
    t0 = time.time()
    inp = case_inputs(kind, dim, q, n, seed)
    model = ref.build_ern(dim, inp["states"])
    clip = ref.FakeClip(inp["text_global"], inp["text_seq"])
    rec0 = ref.Recorder(model)
    if hooks:
        rec0.hook_combiners()  # hooks live on the model; results read via rec0 below

    # pass 1: random targets, only to obtain the reference's ranking
    ref_idx = inp["ref_idx"].clone()
    tgt = inp["rand_tgt"].clone()
    if kind == "cirr":
        same = tgt == ref_idx
        tgt[same] = (tgt[same] + 1) % n
    mem1 = cirr_members(seed + 7, q, n, ref_idx, tgt) if kind == "cirr" else None
    try:
        _, pred1, _, sorted1, *_ = run_reference(kind, dim, inp, ref_idx, tgt, mem1, model, clip)
    except AssertionError:
        raise

    # plant targets at chosen ranks of the reference ranking
    ranks = syn.planted_ranks(seed + 5, q, max_rank=min(100, n - 2))
    planted = torch.empty(q, dtype=torch.long)
    for i in range(q):
        row = sorted1[i]
        if kind == "cirr":
            row = row[row != ref_idx[i]]
        planted[i] = row[int(ranks[i])]
    mem2 = cirr_members(seed + 8, q, n, ref_idx, planted) if kind == "cirr" else None
    if hooks:
        rec0.combiner_io.clear()
    res, pred, rec, sorted2, ref_names, tgt_names, members = run_reference(
        kind, dim, inp, ref_idx, planted, mem2, model, clip)
    assert torch.equal(pred, pred1) and torch.equal(sorted1, sorted2)
    gallery = rec.index_out

    # --- pin the restatement to the reference -------------------------------------------------
    if kind in ("fiq", "shoes"):
        mine = orc.fiq_metrics(pred, gallery, inp["names"], tgt_names, (10, 50))
    elif kind == "val":
        mine = orc.fiq_metrics(pred, gallery, inp["names"], tgt_names, (1, 5, 10, 15, 20, 30, 40, 50))
    elif kind == "200k":
        mine = orc.f200k_metrics(pred, gallery, inp["names"], tgt_names, (10, 50))
    else:
        mine = orc.cirr_metrics(pred, gallery, inp["names"], ref_names, tgt_names, members)
    tie_queries = 0
    if tuple(mine) != tuple(res):
        # The reference's argsort is unstable: inside an EXACT tie of its fp32 distances the order is arbitrary
        # (SURVEY.md P4).  A recall tuple may differ from the stable-sort restatement only through such ties:
        # prove it query by query on the reference's own ranking.
        assert save is False, f"{name}: committed fixtures must be tie-free ({mine} != {res})"
        dfull = orc.distances(pred, gallery)
        for i in range(q):
            row = sorted2[i]
            if kind == "cirr":
                row = row[row != ref_idx[i]]
            r_ref = int((row == planted[i]).nonzero()[0])
            stable = torch.sort(dfull[i], stable=True).indices
            if kind == "cirr":
                stable = stable[stable != ref_idx[i]]
            r_mine = int((stable == planted[i]).nonzero()[0])
            if r_ref != r_mine:
                lo, hi = min(r_ref, r_mine), max(r_ref, r_mine)
                assert float(dfull[i, row[lo]]) == float(dfull[i, row[hi]]), f"{name}: query {i} differs outside a tie"
                tie_queries += 1
        assert tie_queries > 0, f"{name}: tuples differ but no tie explains it"
        assert all(abs(a - b) <= 100.0 * tie_queries / q + 1e-9 for a, b in zip(mine, res))
    ids_o, dist_o = orc.rank_topk(pred, gallery, TOPC)
    dist_full = orc.distances(pred, gallery)
    ref_top = sorted2[:, :TOPC]
    diff = ids_o != ref_top
    # any disagreement must sit inside an exact tie of the reference's own distances
    d_ref = torch.gather(dist_full, 1, ref_top)
    assert torch.equal(d_ref, dist_o), f"{name}: ranked distances differ"
    n_tie = int(diff.sum())
    out = {
        "kind": kind, "dim": dim, "q": q, "n": n, "seed": seed,
        "recall": [float(x) for x in res], "restatement": [float(x) for x in mine], "tie_positions": n_tie,
        "tie_explained_queries": tie_queries,
        "seconds": round(time.time() - t0, 1),
    }
    if save:
        sr = rec0.sr_out if hooks else None
        arrays = dict(
            pred=pred.numpy(), gallery=gallery.numpy(), sr_out=sr.numpy(),
            ref_top=ref_top.numpy().astype(np.int32), ref_dist=d_ref.numpy(),
            ref_idx=ref_idx.numpy().astype(np.int32), tgt_idx=planted.numpy().astype(np.int32),
            members=np.array(mem2 if mem2 is not None else [], dtype=np.int32),
            recall=np.array(res, dtype=np.float64),
            meta=np.array(json.dumps({**out, "digest_index_features": syn.tensor_digest(inp["index_features"]),
                                      "digest_index_local": syn.tensor_digest(inp["index_local"]),
                                      "combiners": list(COMBINERS)})),
        )
        if hooks and name.startswith("fiq640"):
            # the three query-side combiner call sites (models/fusion_model.py:52-54), first 32 rows
            for cname in COMBINERS[:3]:
                io = rec0.combiner_io[cname]
                arrays[f"io_{cname}_image"] = torch.cat([x[0] for x in io])[:32].numpy()
                arrays[f"io_{cname}_text"] = torch.cat([x[1] for x in io])[:32].numpy()
                arrays[f"io_{cname}_out"] = torch.cat([x[2] for x in io])[:32].numpy()
        np.savez(os.path.join(GOLDEN, f"{name}.npz"), **arrays)
    return out


def make_combiner_golden(dim, rows, seed):
    sd = syn.combiner_state(seed, dim)
    m = ref.reference_combiner(dim, sd)
    cases = {}
    # (image, text) as they occur at the four call sites: raw CLIP globals, unit-norm inputs, mixed
    img = syn.features(seed + 1, rows, dim)
    txt = syn.features(seed + 2, rows, dim)
    with torch.no_grad():
        cases["raw"] = (img, txt, m(img, txt))
        iu, tu = syn.features(seed + 3, rows, dim, unit=True), syn.features(seed + 4, rows, dim, unit=True)
        cases["unit"] = (iu, tu, m(iu, tu))
        z = torch.zeros(4, dim)
        cases["zero"] = (z, z, m(z, z))
    for k, (i, t, o) in cases.items():
        mine = orc.combiner_forward(sd, i, t)
        assert torch.equal(mine, o), f"combiner restatement differs ({k})"
    np.savez(os.path.join(GOLDEN, f"combiner{dim}.npz"),
             meta=np.array(json.dumps({"dim": dim, "rows": rows, "seed": seed})),
             **{f"{k}_{nm}": v.numpy() for k, trip in cases.items() for nm, v in zip(("image", "text", "out"), trip)})
    return {"dim": dim, "rows": rows, "seed": seed}


def make_visualsr_golden(dim, rows, seed):
    ref.install()
    from models.fusion_model import VisualSR
    sd = syn.visualsr_state(seed, dim)
    m = VisualSR(embed_dim=dim)
    m.load_state_dict(sd)
    m = m.eval().float()
    x = syn.patch_features(seed + 1, rows, dim)
    with torch.no_grad():
        out = m(x)
    mine = orc.visual_sr_forward(sd, x)
    assert torch.allclose(mine, out, atol=2e-7, rtol=0), float((mine - out).abs().max())
    np.savez(os.path.join(GOLDEN, f"visualsr{dim}.npz"), out=out.numpy(),
             meta=np.array(json.dumps({"dim": dim, "rows": rows, "seed": seed,
                                       "max_abs_diff_restatement": float((mine - out).abs().max())})))
    return {"dim": dim, "rows": rows, "seed": seed}


def make_loss_golden(dim, rows, seed):
    """Reference ``BatchBasedClassificationLoss`` (losses/loss.py) + torch autograd vs the explicit restatement."""
    ref.install()
    from losses.loss import BatchBasedClassificationLoss
    pred, tar = syn.loss_pair(seed, rows, dim)
    p = pred.clone().requires_grad_(True)
    t = tar.clone().requires_grad_(True)
    loss = BatchBasedClassificationLoss()(p, t)
    loss.backward()
    mine, _, dp, dt = orc.bbc_loss(pred, tar)
    err_l = abs(mine - float(loss)) / abs(float(loss))
    err_g = max(float(np.abs(dp - p.grad.numpy()).max()), float(np.abs(dt - t.grad.numpy()).max()))
    assert err_l < 2e-6 and err_g < 2e-6, (err_l, err_g)
    np.savez(os.path.join(GOLDEN, f"bbcloss{dim}.npz"), loss=np.float32(float(loss)), dpred=p.grad.numpy(),
             dtar=t.grad.numpy(),
             meta=np.array(json.dumps({"dim": dim, "rows": rows, "seed": seed, "rel_err_loss_restatement": err_l,
                                       "max_abs_err_grad_restatement": err_g})))
    return {"dim": dim, "rows": rows, "seed": seed, "loss": float(loss), "rel_err_loss_restatement": err_l,
            "max_abs_err_grad_restatement": err_g}


def make_dvr_golden(dim, rows, seed):
    """Reference ``DVR_module`` (BERT + MHA + VisualSR + 3 heads) verbatim vs the explicit restatement."""
    ref.install()
    from models.fusion_model import DVR_module
    sd = syn.dvr_full_state(seed, dim)
    m = DVR_module(feature_dim=dim, device="cpu")
    missing = m.load_state_dict(sd, strict=True)
    m = m.eval().float()
    patches = syn.patch_features(seed + 10, rows, dim)
    tokens = syn.token_features(seed + 11, rows, dim)
    ref_g, txt_g = syn.features(seed + 12, rows, dim), syn.features(seed + 13, rows, dim)
    with torch.no_grad():
        out = m(patches, tokens, ref_g, txt_g)
        _, hidden, _ = m.transformer_layer(patches, tokens)
    mine, hid = orc.dvr_forward(sd, patches, tokens, ref_g, txt_g, return_hidden=True)
    err_h, err_o = float((hid - hidden).abs().max()), float((mine - out).abs().max())
    assert err_h < 2e-5 and err_o < 2e-6, (err_h, err_o)
    np.savez(os.path.join(GOLDEN, f"dvr{dim}.npz"), out=out.numpy(), hidden_cls=hidden[:, 0].numpy(),
             hidden_last=hidden[:, -1].numpy(),
             meta=np.array(json.dumps({"dim": dim, "rows": rows, "seed": seed, "err_hidden": err_h, "err_out": err_o})))
    return {"dim": dim, "rows": rows, "seed": seed, "err_hidden_restatement": err_h, "err_out_restatement": err_o}


def make_full_model_golden(name, kind, dim, q, n, seed):
    """The WHOLE reference model (BERT, MHA, VisualSR, four heads) with synthetic weights through the unmodified
    ``compute_*_val_metrics``: pins the fully accelerated ERN (DVR + SR + heads on the B200 path) end to end."""
    inp = case_inputs(kind, dim, q, n, seed)
    ref.install()
    from models.model import ERN
    model = ERN(ref.FakeClip(None, None), dim, "cpu")
    model.load_state_dict(syn.ern_full_state(seed + 50, dim))
    model = model.eval().float()
    clip = ref.FakeClip(inp["text_global"], inp["text_seq"])
    ref_idx, tgt = inp["ref_idx"].clone(), inp["rand_tgt"].clone()
    same = tgt == ref_idx
    tgt[same] = (tgt[same] + 1) % n
    mem = cirr_members(seed + 7, q, n, ref_idx, tgt) if kind == "cirr" else None
    _, pred1, _, sorted1, *_ = run_reference(kind, dim, inp, ref_idx, tgt, mem, model, clip)
    ranks = syn.planted_ranks(seed + 5, q, max_rank=min(100, n - 2))
    planted = torch.empty(q, dtype=torch.long)
    for i in range(q):
        row = sorted1[i]
        if kind == "cirr":
            row = row[row != ref_idx[i]]
        planted[i] = row[int(ranks[i])]
    mem2 = cirr_members(seed + 8, q, n, ref_idx, planted) if kind == "cirr" else None
    res, pred, rec, sorted2, ref_names, tgt_names, members = run_reference(kind, dim, inp, ref_idx, planted, mem2, model, clip)
    gallery = rec.index_out
    d_ref = torch.gather(orc.distances(pred, gallery), 1, sorted2[:, :TOPC])
    np.savez(os.path.join(GOLDEN, f"{name}.npz"), pred=pred.numpy(), gallery_head=gallery[:16].numpy(),
             ref_top=sorted2[:, :TOPC].numpy().astype(np.int32), ref_dist=d_ref.numpy(),
             ref_idx=ref_idx.numpy().astype(np.int32), tgt_idx=planted.numpy().astype(np.int32),
             members=np.array(mem2 if mem2 is not None else [], dtype=np.int32), recall=np.array(res, dtype=np.float64),
             meta=np.array(json.dumps({"kind": kind, "dim": dim, "q": q, "n": n, "seed": seed, "recall": list(map(float, res)),
                                       "digest_index_features": syn.tensor_digest(inp["index_features"]),
                                       "digest_index_local": syn.tensor_digest(inp["index_local"])})))
    return {"kind": kind, "dim": dim, "q": q, "n": n, "seed": seed, "recall": [float(x) for x in res]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-full", action="store_true")
    ap.add_argument("--only-full", action="store_true", help="keep the committed fixtures, redo the full-size pins")
    ap.add_argument("--only-visualsr", action="store_true", help="only (re)make the VisualSR fixtures")
    ap.add_argument("--only-loss", action="store_true", help="only (re)make the training-criterion fixtures")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    loss_cases = [make_loss_golden(640, 48, 700), make_loss_golden(512, 100, 710)]
    if args.only_loss:
        with open(os.path.join(GOLDEN, "pin_report.json")) as f:
            old = json.load(f)
        old["bbc_loss"] = loss_cases
        with open(os.path.join(GOLDEN, "pin_report.json"), "w") as f:
            json.dump(old, f, indent=1)
        print(loss_cases)
        return
    report = {"torch": torch.__version__, "numpy": np.__version__, "cases": {}, "full": {}}
    report["bbc_loss"] = loss_cases
    report["visualsr"] = [make_visualsr_golden(640, 40, 300), make_visualsr_golden(512, 40, 400)]
    report["dvr"] = [make_dvr_golden(640, 6, 500), make_dvr_golden(512, 6, 600)]
    report["full_model"] = [make_full_model_golden("ernfull_fiq640", "fiq", 640, 40, 160, 1300),
                            make_full_model_golden("ernfull_cirr512", "cirr", 512, 40, 160, 1310)]
    if args.only_visualsr:
        with open(os.path.join(GOLDEN, "pin_report.json")) as f:
            old = json.load(f)
        old["visualsr"] = report["visualsr"]
        old["dvr"] = report["dvr"]
        old["full_model"] = report["full_model"]
        with open(os.path.join(GOLDEN, "pin_report.json"), "w") as f:
            json.dump(old, f, indent=1)
        return
    if args.only_full:
        with open(os.path.join(GOLDEN, "pin_report.json")) as f:
            old = json.load(f)
        old["visualsr"] = report["visualsr"]
        old["dvr"] = report["dvr"]
        old["full_model"] = report["full_model"]
        report = old
        report["full"] = {}
    else:
        report["combiner"] = [make_combiner_golden(640, 48, 100), make_combiner_golden(512, 48, 200)]
    small = [
        ("fiq640", "fiq", 640, 48, 192, 1234),
        ("val512", "val", 512, 48, 192, 1240),
        ("shoes640", "shoes", 640, 40, 160, 1250),
        ("f200k640", "200k", 640, 48, 192, 1260),
        ("cirr640", "cirr", 640, 48, 160, 1270),
    ]
    for name, kind, dim, q, n, seed in ([] if args.only_full else small):
        report["cases"][name] = make_case(name, kind, dim, q, n, seed)
        print(name, report["cases"][name], flush=True)
    if not args.no_full:
        # full dataset shapes: pins only, nothing large is stored
        for name, kind, dim, q, n, seed in [("fiq_dress_full", "fiq", 640, 2017, 3817, 1234),
                                            ("cirr_val_full", "cirr", 640, 4181, 2297, 1270)]:
            report["full"][name] = make_case(name, kind, dim, q, n, seed, save=False, hooks=False)
            print(name, report["full"][name], flush=True)
    with open(os.path.join(GOLDEN, "pin_report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
