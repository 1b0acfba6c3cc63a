"""GPU: the tail at the dataset shapes BASELINE.json names (configs 2-4), against the CPU oracle on the same
seeded synthetic features.  fp32 validation mode must reproduce the oracle's Recall tuples bit for bit; bf16
mode is checked rank-wise against the oracle fed the same bf16-rounded operands."""
import numpy as np
import pytest
import torch

import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import metrics, ops, synthetic as syn
from fashionern_aaai2024_b200._lib import MODE_BF16, RANK_REFERENCE
from oracle import ern_oracle as orc

pytestmark = pytest.mark.gpu


def planted(pred, gal, seed, kmax, exclude=None):
    ids_o, _ = orc.rank_topk(pred, gal, kmax + 1, exclude_index=exclude)
    q = pred.shape[0]
    return ids_o, ids_o[torch.arange(q), syn.planted_ranks(seed, q, kmax + 1).clamp(max=kmax)]


@pytest.mark.parametrize("q,n,seed", [(2017, 3817, 1235), (2038, 6346, 1236), (1961, 5373, 1237)])
def test_config2_fashioniq_categories_512d(cuda_device, q, n, seed):
    dim = 512
    pred, gal = syn.features(seed, q, dim, unit=True), syn.features(seed + 100, n, dim, unit=True)
    names = syn.unique_names(n)
    _, tgt = planted(pred, gal, seed + 200, 50)
    tgt_names = [names[int(i)] for i in tgt]
    want = orc.fiq_metrics(pred, gal, names, tgt_names, (10, 50))
    got = metrics._unique_tail(pred.to(cuda_device), gal.to(cuda_device), names, tgt_names, (10, 50), "fp32", cuda_device)
    assert got == want
    # bf16 product path: identical ranking to the oracle on the same rounded operands (outside 2e-6 near-ties)
    qb, gb = pred.bfloat16(), gal.bfloat16()
    vals, ids, _, _ = ops.sim_topk(qb.to(cuda_device), gb.to(cuda_device), 50, mode=MODE_BF16, rank_by=RANK_REFERENCE)
    st = orc.compare_topk(ids.cpu().numpy(), None, qb.float(), gb.float(), 50, tol=2e-6)
    assert st["exact_frac"] > 0.995


def test_config3_fashion200k_shape_anyhit(cuda_device):
    q, n, dim, classes = 33480, 29789, 640, 5000
    pred, gal = syn.features(3001, q, dim, unit=True), syn.features(3002, n, dim, unit=True)
    names = syn.caption_names(3003, n, classes)                       # non-unique caption "names"
    gcls, = orc.factorize(names)
    ids_o, d_o = orc.rank_topk(pred, gal, 51)
    tgt_rows = ids_o[torch.arange(q), syn.planted_ranks(3004, q, 50).clamp(max=49)]
    ids_o = ids_o[:, :50]
    tgt_names = [names[int(i)] for i in tgt_rows]
    _, tcls = orc.factorize(names, tgt_names)
    want = orc.recall_at(orc.first_hit_rank(ids_o, gcls, tcls), (1, 10, 50))
    res = ern.score_topk_recall(pred.to(cuda_device), gal.to(cuda_device),
                                torch.from_numpy(gcls.astype(np.int32)).to(cuda_device),
                                torch.from_numpy(tcls.astype(np.int32)).to(cuda_device), (1, 10, 50), precision="fp32")
    # fp32 sums in a different order: a query may cross a K boundary only where the oracle's own distances
    # at ranks K-1 and K are closer than 2e-6
    for got, exp, kk in zip(res["recall"], want, (1, 10, 50)):
        near = int((d_o[:, kk] - d_o[:, kk - 1] < 2e-6).sum())
        assert abs(got - exp) <= 100.0 * near / q + 1e-9, (kk, got, exp, near)
    assert 5 < want[0] < want[1] < want[2] <= 100
    # bf16 product path at the full 33 480 x 29 789 shape: exact against the oracle on the same bf16-rounded operands
    from helpers import assert_tuple_within_ties, rounded, unique_oracle_with_ties
    pred_r, gal_r = rounded(pred), rounded(gal)
    want_b, near_b = unique_oracle_with_ties(pred_r, gal_r, names, tgt_names, (1, 10, 50), anyhit=True)
    res_b = ern.score_topk_recall(pred.to(cuda_device), gal.to(cuda_device),
                                  torch.from_numpy(gcls.astype(np.int32)).to(cuda_device),
                                  torch.from_numpy(tcls.astype(np.int32)).to(cuda_device), (1, 10, 50), precision="bf16")
    assert_tuple_within_ties(res_b["recall"], want_b, near_b, q)


def test_config4_cirr_shape(cuda_device):
    q, n, dim = 4181, 2297, 640
    pred, gal = syn.features(4001, q, dim, unit=True), syn.features(4002, n, dim, unit=True)
    names = syn.unique_names(n, "dev-{}-img")
    g = torch.Generator().manual_seed(4003)
    ref = torch.randint(0, n, (q,), generator=g)
    _, tgt = planted(pred, gal, 4004, 50, exclude=ref)
    members = []
    for i in range(q):
        r, t = int(ref[i]), int(tgt[i])
        others = [x for x in torch.randint(0, n, (12,), generator=g).tolist() if x != r and x != t]
        others = list(dict.fromkeys(others))[:4]
        row = [r, t] + others
        perm = torch.randperm(len(row), generator=g).tolist()
        members.append([names[row[p]] for p in perm] + [names[r]] * (6 - len(row)))
    ref_names = [names[int(i)] for i in ref]
    tgt_names = [names[int(i)] for i in tgt]
    want = orc.cirr_metrics(pred, gal, names, ref_names, tgt_names, members)
    got = metrics.cirr_tail(pred.to(cuda_device), gal.to(cuda_device), names, ref_names, tgt_names, members, "fp32", cuda_device)
    assert got == want
    # bf16 product path: R@K and the subset ranks G@K exact against the oracle on the same bf16-rounded operands
    from helpers import assert_tuple_within_ties, cirr_oracle_with_ties, rounded
    got_b = metrics.cirr_tail(pred.to(cuda_device), gal.to(cuda_device), names, ref_names, tgt_names, members, "bf16", cuda_device)
    want_b, near_b = cirr_oracle_with_ties(rounded(pred), rounded(gal), names, ref_names, tgt_names, members)
    assert_tuple_within_ties(got_b, want_b, near_b, q)
    assert all(abs(a - b) < 3.0 for a, b in zip(got_b, want))          # and close to the fp32 tuple


def test_scale_properties_10m(cuda_device):
    """Size-independent properties at a gallery too large for the CPU oracle (10M x 640):
    (i) sharding invariance: merged shard results == unsharded result (keys bit-identical);
    (ii) planted exact duplicates of a query are returned at rank 0 with similarity ~1;
    (iii) rows are sorted and duplicate-free; (iv) schedule invariance (growth 8 vs 4)."""
    q, n, dim, k = 512, 10_000_000, 640, 100
    gen = torch.Generator(device=cuda_device).manual_seed(77)
    gal = torch.empty(n, dim, dtype=torch.bfloat16, device=cuda_device)
    for s in range(0, n, 1 << 20):
        x = torch.randn(min(1 << 20, n - s), dim, generator=gen, device=cuda_device)
        gal[s:s + (1 << 20)] = torch.nn.functional.normalize(x, dim=-1).bfloat16()
    rows = torch.randint(0, n, (q,), generator=torch.Generator().manual_seed(78)).to(cuda_device)
    pred = gal[rows].clone()                                                     # queries = planted gallery rows
    vals, ids, keys, status = ops.sim_topk(pred, gal, k, want_keys=True)
    assert int(status[0].item()) == 0
    hit0 = (ids[:, 0].long() == rows) | (vals[:, 0] == vals[:, 1])               # exact duplicate rows may tie
    assert bool(hit0.all()) and bool((vals[:, 0] > 0.98).all())
    assert bool((vals[:, :-1] >= vals[:, 1:]).all())
    srt = ids.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    bounds = [0, 3_000_000, 3_000_001, 7_654_321, n]
    parts = [ops.sim_topk(pred, gal[a:b], k, id_offset=a, want_keys=True)[2] for a, b in zip(bounds[:-1], bounds[1:])]
    _, ids_m, keys_m = ops.topk_merge(torch.stack(parts), k)
    assert torch.equal(keys_m, keys) and torch.equal(ids_m, ids)
    assert torch.equal(ops.sim_topk(pred, gal, k, growth=4, want_keys=True)[2], keys)
