"""GPU: similarity + top-k (ern_sim_topk), merge, recall kernels against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import ern_oracle as orc
from fashionern_aaai2024_b200 import ops, synthetic as syn
from fashionern_aaai2024_b200._lib import MODE_BF16, MODE_FP32, RANK_REFERENCE, RANK_SIMILARITY

pytestmark = pytest.mark.gpu

# Tolerances: the CUDA and oracle scores are fp32 sums of the same products in a
# different order, so they agree to a few ulp of the largest partial sums; ids must be identical except
# where the ORACLE's own scores are closer than this.
TOL_FP32 = 2e-6
TOL_BF16_SAME_INPUTS = 2e-6   # oracle fed the same bf16-rounded operands: only accumulation order differs
# Secondary check next to compare_topk's rank-wise gap rule: the fraction of positions that are IDENTICAL to the
# oracle's.  Observed minimum over every parametrised case below (tools/exact_frac_census.py,
# profiles/r02_exact_frac_census.jsonl): 1.0 in fp32, 0.9996 with bf16 and 0.9998 with fp16 operands -- the rest are
# swaps inside < 2e-6 near-ties.
EXACT_FRAC_16BIT = 0.999


def unit(seed, rows, dim):
    return syn.features(seed, rows, dim, unit=True)


def run_case(dev, q, n, dim, k, mode, rank_by=RANK_REFERENCE, exclude=None, growth=8, seed=0, cast=torch.bfloat16):
    pred, gal = unit(seed + 1, q, dim), unit(seed + 2, n, dim)
    if mode == MODE_BF16:       # the tensor-core mode: bf16 operands, or fp16 (cast=torch.float16)
        pred_o, gal_o = pred.to(cast).float(), gal.to(cast).float()
        qd, gd = pred.to(cast).to(dev), gal.to(cast).to(dev)
        tol = TOL_BF16_SAME_INPUTS
    else:
        pred_o, gal_o = pred, gal
        qd, gd = pred.to(dev), gal.to(dev)
        tol = TOL_FP32
    ex = None if exclude is None else exclude.to(dev)
    vals, ids, keys, status = ops.sim_topk(qd, gd, k, mode=mode, rank_by=rank_by, exclude_ids=ex, growth=growth,
                                           want_keys=True)
    torch.cuda.synchronize()
    ids_c, vals_c = ids.cpu().numpy(), vals.cpu().numpy()
    sims = vals_c if rank_by == RANK_SIMILARITY else vals_c + 1.0   # -(1-s) -> s (approximately)
    finite = np.where(np.isfinite(sims), sims, 0.0)
    stats = orc.compare_topk(ids_c, finite, pred_o, gal_o, k, tol=tol + (0 if rank_by == RANK_SIMILARITY else 1.2e-7),
                             exclude_index=exclude)
    # rows are sorted best-first, ties by lower id
    v = vals_c
    assert np.all((v[:, :-1] >= v[:, 1:]) | ~np.isfinite(v[:, 1:]))
    tie = (v[:, :-1] == v[:, 1:]) & np.isfinite(v[:, 1:])
    assert np.all(ids_c[:, :-1][tie] < ids_c[:, 1:][tie])
    return stats, ids, vals, keys


@pytest.mark.parametrize("q,n,dim,k", [(48, 192, 640, 50), (1, 1, 64, 5), (7, 33, 100, 10), (130, 3000, 512, 51),
                                       (257, 5000, 640, 100), (64, 2048, 64, 128), (65, 2049, 64, 128), (33, 256, 64, 128), (33, 257, 64, 128)])
def test_fp32_validation_mode_matches_oracle(cuda_device, q, n, dim, k):
    stats, *_ = run_case(cuda_device, q, n, dim, k, MODE_FP32)
    assert stats["exact_frac"] > 0.999


@pytest.mark.parametrize("q,n,dim,k", [(48, 192, 640, 50), (100, 5000, 640, 100), (128, 2300, 512, 50),
                                       (5, 130, 64, 10), (1, 4000, 640, 1), (100, 9000, 768, 50), (64, 3000, 704, 20)])
def test_bf16_single_cta_matches_oracle(cuda_device, q, n, dim, k):
    stats, *_ = run_case(cuda_device, q, n, dim, k, MODE_BF16)   # q <= 128 -> 1-CTA tcgen05 kernel
    assert stats["exact_frac"] > EXACT_FRAC_16BIT


@pytest.mark.parametrize("q,n,dim,k", [(300, 5000, 640, 100), (129, 257, 640, 50), (2017, 3817, 640, 51),
                                       (512, 40000, 512, 100), (1000, 20000, 128, 128), (700, 30000, 768, 100)])
def test_bf16_cta_pair_matches_oracle(cuda_device, q, n, dim, k):
    stats, *_ = run_case(cuda_device, q, n, dim, k, MODE_BF16)   # q > 128 -> cta_group::2 kernel
    assert stats["exact_frac"] > EXACT_FRAC_16BIT


@pytest.mark.parametrize("q,n,dim,k", [(100, 5000, 640, 100), (5, 130, 64, 10), (300, 5000, 640, 100),
                                       (2017, 3817, 640, 51), (700, 30000, 768, 100)])
def test_fp16_operands_match_oracle(cuda_device, q, n, dim, k):
    # ERN_DTYPE_F16: the same tcgen05 kernels (1-CTA and CTA pair) with the fp16 operand format; the oracle gets the
    # same fp16-rounded operands, so only the accumulation order differs
    stats, *_ = run_case(cuda_device, q, n, dim, k, MODE_BF16, cast=torch.float16)
    assert stats["exact_frac"] > EXACT_FRAC_16BIT


def test_fp16_operands_are_closer_to_fp32_than_bf16(cuda_device):
    # unit-norm 640-d features: fp16 keeps 11 mantissa bits (bf16: 8), and nothing is near its range limits
    q, n, dim, k = 256, 20000, 640, 100
    pred, gal = unit(11, q, dim), unit(12, n, dim)
    v32, i32, _, _ = ops.sim_topk(pred.to(cuda_device), gal.to(cuda_device), k, mode=MODE_FP32)
    err = {}
    for cast in (torch.float16, torch.bfloat16):
        qd, gd = pred.to(cast).to(cuda_device), gal.to(cast).to(cuda_device)
        assert ops.gather_scores(qd, gd, i32[:, :50]).sub(v32[:, :50]).abs().max().item() < (2e-4 if cast == torch.float16 else 2e-3)
        v, _, _, _ = ops.sim_topk(qd, gd, k)
        err[cast] = (v - v32).abs().max().item()        # k-th best values side by side
    assert err[torch.float16] < 2e-4 and err[torch.float16] < err[torch.bfloat16]
    with pytest.raises(ops.ErnError):
        ops.sim_topk(pred.half().to(cuda_device), gal.bfloat16().to(cuda_device), k)       # mixed 16-bit types
    with pytest.raises(ops.ErnError):
        ops.sim_topk(pred.half().to(cuda_device), gal.half().to(cuda_device), k, mode=MODE_FP32)


def test_bf16_similarity_ranking_and_exclusion(cuda_device):
    q, n = 200, 6000
    ex = torch.randint(0, n, (q,), generator=torch.Generator().manual_seed(5))
    ex[::7] = -1
    run_case(cuda_device, q, n, 640, 50, MODE_BF16, rank_by=RANK_SIMILARITY, exclude=ex)
    run_case(cuda_device, q, n, 640, 50, MODE_FP32, rank_by=RANK_REFERENCE, exclude=ex)


def test_fewer_rows_than_k_pads(cuda_device):
    pred, gal = unit(1, 9, 64).to(cuda_device), unit(2, 5, 64).to(cuda_device)
    vals, ids, _, _ = ops.sim_topk(pred, gal, 8, mode=MODE_FP32)
    assert bool((ids[:, 5:] == -1).all()) and bool(torch.isinf(vals[:, 5:]).all())
    assert bool((ids[:, :5].sort(dim=1).values == torch.arange(5, device=cuda_device, dtype=torch.int32)).all())
    vals, ids, _, _ = ops.sim_topk(pred.bfloat16(), gal.bfloat16(), 8, mode=MODE_BF16)
    assert bool((ids[:, 5:] == -1).all())
    v0, i0, _, _ = ops.sim_topk(pred, gal[:0], 4, mode=MODE_FP32)        # empty gallery
    assert bool((i0 == -1).all())
    v1, i1, _, _ = ops.sim_topk(pred[:0], gal, 4, mode=MODE_FP32)        # empty query batch
    assert i1.shape == (0, 4)


def test_duplicate_rows_tie_break_lowest_id(cuda_device):
    g = unit(3, 40, 64)
    gal = torch.cat([g, g, g])                     # every row three times -> exact ties
    pred = unit(4, 6, 64)
    for mode, cast in ((MODE_FP32, torch.float32), (MODE_BF16, torch.bfloat16)):
        vals, ids, _, _ = ops.sim_topk(pred.to(cuda_device, cast), gal.to(cuda_device, cast), 9, mode=mode)
        ids = ids.cpu().numpy()
        assert np.all(ids[:, 0::3] + 40 == ids[:, 1::3]) and np.all(ids[:, 1::3] + 40 == ids[:, 2::3])


def test_adversarial_order_needs_no_fallback(cuda_device):
    # gallery sorted by increasing similarity to the query: every row beats the running threshold.  The candidate
    # segments compact themselves inside the kernel (fp32 mode: launches never exceed a query's candidate slots), so
    # the very first call is exact -- there is no overflow status and no conservative re-run any more
    # (more orders, batch sizes and k in tests/test_gpu_order.py)
    n, dim = 30000, 64
    base = unit(7, 1, dim)
    noise = unit(8, n, dim)
    w = torch.linspace(0.0, 1.0, n)[:, None]
    gal = torch.nn.functional.normalize(w * base + (1 - w) * 0.3 * noise, dim=-1)
    pred = base.repeat(3, 1)
    qd, gd = pred.to(cuda_device), gal.to(cuda_device)
    vals, ids, _, status = ops.sim_topk(qd, gd, 100, mode=MODE_FP32, check_overflow=False)
    assert status.cpu().tolist() == [0, 0, 0, 0]
    orc.compare_topk(ids.cpu().numpy(), None, pred, gal, 100, tol=TOL_FP32 + 1.2e-7)
    _, ids_b, _, status = ops.sim_topk(qd.bfloat16(), gd.bfloat16(), 100, mode=MODE_BF16, check_overflow=False)
    assert status.cpu().tolist() == [0, 0, 0, 0]
    orc.compare_topk(ids_b.cpu().numpy(), None, pred.bfloat16().float(), gal.bfloat16().float(), 100, tol=2e-6)


def test_growth_schedules_agree(cuda_device):
    q, n, k = 150, 50000, 100
    pred, gal = unit(11, q, 128).bfloat16().to(cuda_device), unit(12, n, 128).bfloat16().to(cuda_device)
    ref = ops.sim_topk(pred, gal, k)[1]
    for growth in (1, 2, 4, 16, 64):
        assert torch.equal(ops.sim_topk(pred, gal, k, growth=growth)[1], ref)


def test_shard_merge_equals_global(cuda_device):
    q, n, k = 140, 9000, 100
    pred, gal = unit(21, q, 640).bfloat16().to(cuda_device), unit(22, n, 640).bfloat16().to(cuda_device)
    _, ids_all, keys_all, _ = ops.sim_topk(pred, gal, k, want_keys=True)
    bounds = [0, 2250, 4500, 6750, 9000]
    parts = [ops.sim_topk(pred, gal[a:b], k, id_offset=a, want_keys=True)[2] for a, b in zip(bounds[:-1], bounds[1:])]
    vals, ids, keys = ops.topk_merge(torch.stack(parts), k)
    assert torch.equal(ids, ids_all) and torch.equal(keys, keys_all)


@pytest.mark.parametrize("q,world", [(140, 3), (100, 2), (600, 4), (4200, 2)])   # 4200: two query batches
def test_fused_exchange_stores_on_one_gpu(cuda_device, q, world):
    # ern_sim_topk_exchange on ONE device: the `world` gathered buffers all live here and the peer-pointer table is
    # built by hand (what torch symmetric memory hands out on a multi-GPU box, tools/test_exchange.py).  Every "rank"
    # scores its shard and its last selection launch stores the keys into slot `rank` of EVERY buffer: each buffer
    # must end up as the stack of the per-shard key lists, bit for bit, and merge to the global top-k.
    n, k, dim = 9001, 100, 640
    pred, gal = unit(23, q, dim).bfloat16().to(cuda_device), unit(24, n, dim).bfloat16().to(cuda_device)
    _, ids_all, keys_all, _ = ops.sim_topk(pred, gal, k, want_keys=True)
    per = (n + world - 1) // world
    bounds = [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]
    parts = torch.stack([ops.sim_topk(pred, gal[a:b], k, id_offset=a, want_keys=True)[2] for a, b in bounds])
    bufs = [torch.full((world, q, k), -7, dtype=torch.int64, device=cuda_device) for _ in range(world)]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=cuda_device)
    for r, (a, b) in enumerate(bounds):
        st = ops.sim_topk_exchange(pred, gal[a:b], k, table.data_ptr(), world, r, id_offset=a)
        assert st.cpu().tolist() == [0, 0, 0, 0]
    for buf in bufs:
        assert torch.equal(buf, parts)
        _, ids, keys = ops.topk_merge(buf, k)
        assert torch.equal(ids, ids_all) and torch.equal(keys, keys_all)


def test_recall_kernels_match_oracle(cuda_device):
    q, n, k = 500, 4000, 50
    pred, gal = unit(31, q, 640), unit(32, n, 640)
    g = torch.Generator().manual_seed(33)
    cls = torch.randint(0, 700, (n,), generator=g)            # non-unique classes (Fashion200k-style)
    ids_o, _ = orc.rank_topk(pred, gal, k)
    tgt_cls = cls[ids_o[torch.arange(q), syn.planted_ranks(34, q, 60).clamp(max=k - 1)]]
    _, ids, _, _ = ops.sim_topk(pred.to(cuda_device), gal.to(cuda_device), k, mode=MODE_FP32)
    counts, ranks = ops.recall_at_k(ids, cls.int().to(cuda_device), tgt_cls.int().to(cuda_device), (1, 10, 50))
    exp_ranks = orc.first_hit_rank(ids.cpu().long(), cls.numpy(), tgt_cls.numpy())
    assert np.array_equal(ranks.cpu().numpy(), exp_ranks)
    assert counts.cpu().tolist() == [int((exp_ranks < kk).sum()) for kk in (1, 10, 50)]


def test_large_gallery_bf16_subset_parity(cuda_device):
    # 4096 x 1M x 640 (SURVEY.md 8d config 5 at its smallest size): parity on a 64-query subset against
    # fp32 matmul + stable sort of the same bf16-rounded data (oracle), gap-aware
    q, n, dim, k = 4096, 1_000_000, 640, 100
    gen = torch.Generator(device=cuda_device).manual_seed(5000)
    gal = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen, device=cuda_device), dim=-1).bfloat16()
    pred = torch.nn.functional.normalize(torch.randn(q, dim, generator=gen, device=cuda_device), dim=-1).bfloat16()
    vals, ids, _, status = ops.sim_topk(pred, gal, k)
    assert int(status[0].item()) == 0
    sub = torch.arange(0, q, 64)
    orc.compare_topk(ids[sub.to(cuda_device)].cpu().numpy(), None, pred[sub.to(cuda_device)].float().cpu(),
                     gal.float().cpu(), k, tol=TOL_BF16_SAME_INPUTS)


def test_split_cirr_subset_equals_fused(cuda_device):
    # shard-emulation on one GPU: member scores gathered shard by shard and summed == the fused kernel
    q, n, dim = 300, 2000, 640
    pred, gal = unit(51, q, dim).to(cuda_device), unit(52, n, dim).to(cuda_device)
    g = torch.Generator().manual_seed(53)
    members = torch.stack([torch.randperm(n, generator=g)[:6] for _ in range(q)]).int().to(cuda_device)
    ref, tgt = members[:, 0].contiguous(), members[:, 1].contiguous()
    members = members[:, torch.randperm(6, generator=g)].contiguous()
    for cast in (torch.float32, torch.bfloat16):
        p, G = pred.to(cast), gal.to(cast)
        c_f, r_f = ops.cirr_subset_recall(p, G, members, ref, tgt, (1, 2, 3))
        total = torch.zeros(q, 6, device=cuda_device)
        for a, b in ((0, 700), (700, 701), (701, 2000)):
            total += ops.gather_scores(p, G[a:b], members, id_offset=a)
        c_s, r_s = ops.cirr_subset_from_scores(total, members, ref, tgt, (1, 2, 3))
        assert torch.equal(r_f, r_s) and torch.equal(c_f, c_s)
        assert bool((r_s >= 0).all()) and 0 < int(c_s[0]) < q


def test_cuda_graph_capture_of_the_tail(cuda_device):
    """Every C-ABI call is stream-ordered and allocation-free, so the whole tail (cast -> scoring/top-k -> recall)
    can be captured once in a CUDA graph and replayed on new data (launch-bound dataset-scale shapes)."""
    q, n, dim, k = 500, 4000, 640, 50
    pred_s, gal_s = torch.empty(q, dim, device=cuda_device), torch.empty(n, dim, device=cuda_device)
    cls = torch.arange(n, dtype=torch.int32, device=cuda_device)
    tgt_s = torch.zeros(q, dtype=torch.int32, device=cuda_device)

    def tail():
        _, qb = ops.l2norm_rows(pred_s, normalize=False, want_f32=False, want_bf16=True)
        _, gb = ops.l2norm_rows(gal_s, normalize=False, want_f32=False, want_bf16=True)
        vals, ids, _, status = ops.sim_topk(qb, gb, k, check_overflow=False)
        counts, ranks = ops.recall_at_k(ids, cls, tgt_s, (1, 10, 50))
        return vals, ids, counts, status

    pred_s.copy_(unit(61, q, dim)); gal_s.copy_(unit(62, n, dim))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            tail()                                   # warm-up outside capture (lazy kernel attributes, entry points)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_vals, g_ids, g_counts, g_status = tail()
    for seed in (63, 65):
        pred_s.copy_(unit(seed, q, dim)); gal_s.copy_(unit(seed + 1, n, dim))
        tgt_s.copy_(torch.randint(0, n, (q,), generator=torch.Generator().manual_seed(seed)).int())
        graph.replay()
        torch.cuda.synchronize()
        e_vals, e_ids, e_counts, _ = tail()
        assert torch.equal(g_ids, e_ids) and torch.equal(g_vals, e_vals) and torch.equal(g_counts, e_counts)
        assert int(g_status[0].item()) == 0


def test_sharded_helpers_single_process(cuda_device):
    # world_size 1 (no process group): the sharded entry points must reduce to the single-GPU calls
    from fashionern_aaai2024_b200 import sharded
    q, n, dim = 130, 3000, 640
    pred, gal = unit(71, q, dim).bfloat16().to(cuda_device), unit(72, n, dim).bfloat16().to(cuda_device)
    cls = torch.arange(n, dtype=torch.int32, device=cuda_device)
    tgt = torch.randint(0, n, (q,), generator=torch.Generator().manual_seed(73)).int().to(cuda_device)
    v0, i0, k0, _ = ops.sim_topk(pred, gal, 50, want_keys=True)
    v1, i1, k1, st = sharded.sharded_topk(pred, gal, 50, 0)
    assert torch.equal(i0, i1) and torch.equal(k0, k1) and int(st[0].item()) == 0
    c0, r0 = ops.recall_at_k(i0, cls, tgt, (1, 10, 50))
    c1, r1, _ = sharded.sharded_recall(pred, gal, 0, cls, tgt, (1, 10, 50))
    assert torch.equal(c0, c1) and torch.equal(r0, r1)
    g = torch.Generator().manual_seed(74)
    members = torch.stack([torch.randperm(n, generator=g)[:6] for _ in range(q)]).int().to(cuda_device)
    ref, tg = members[:, 0].contiguous(), members[:, 1].contiguous()
    a = ops.cirr_subset_recall(pred, gal, members, ref, tg, (1, 2, 3))
    b = sharded.sharded_cirr_subset(pred, gal, 0, members, ref, tg, (1, 2, 3))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
