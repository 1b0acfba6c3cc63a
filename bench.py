#!/usr/bin/env python
"""Headline benchmark of the composed-retrieval scoring path (BASELINE.json: queries/sec, top-100, 640-d,
synthetic gallery, 4096-query batches, gallery row-sharded over N B200s of one box).

    python bench.py --gpus 1 --steps 5 --warmup 3                       # ours, N = 1 (default: 100M rows)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                            # ours, N ranks
    python bench.py --impl reference ...                                  # the reference's CPU path
    python bench.py --gallery-rows 10000000 | --gallery-order clustered   # BASELINE configs[4] sweep / ordered gallery
    python bench.py --config fiq|fiq512-dress|fiq512-shirt|fiq512-toptee|shoes|f200k|cirr   # dataset shapes (configs 1-4)

One STEP of the default (scaling) config = one batch of 4096 composed queries through the whole hot path:
    fusion head (CombinerSimple: image+text CLIP features -> unit-norm query, tcgen05 GEMMs)
 -> bf16 cosine scoring of the batch against this rank's gallery shard with streaming top-100 (tcgen05)
 -> [N > 1] exchange of the (score,id) candidate keys (fused peer-memory stores over NVLink, or NCCL all-gather
    with --exchange nccl) + device k-way merge
 -> Recall@{1,10,50,100} hit counts from id membership on device.
One STEP of a dataset config = the metric tail of `compute_*_val_metrics` at that dataset's shape: gallery
L2-normalise -> model(mode="index") (VisualSR + fusion head over the whole gallery) -> scoring + top-50 ->
Recall@K (CIRR: reference removal + subset recall), run/test/test_fiq.py:44-64 and twins.

`value`  : queries/s of the whole job with every input already resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public API with HOST inputs: per step the query batch's features and target ids
           are copied from pinned host memory, and the top-k ids/scores + hit counts are read back.
`parity_check` : BEFORE timing, the top-k this very run produces is verified against an independent reference on the
           same bf16 operands -- every (query, gallery row) score is recomputed with a cuBLAS bf16->fp32 GEMM
           (`torch.mm(out_dtype=float32)`), all rows of every rank's shard, and the returned set / values / order are
           checked to be the exact top-k up to score gaps <= 2e-6; a 16-query subset is additionally re-ranked with
           an fp32 (non tensor-core) matmul + torch.topk.  Targets for Recall@K are planted from the verified ranking.
The gallery (rows x 640 bf16 = 128 GB at 100M rows) is far larger than L2 (126 MB), so every step streams it
from HBM: no L2 flush is needed between timed iterations (stated in `config.l2`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KS = (1, 10, 50, 100)
TOL = 2e-6                      # score gap below which two gallery rows may swap places (BASELINE.md tolerance)
REFERENCE_ROOT = "/root/reference"

# dataset shapes of BASELINE.json configs[0..3] (public val/test split sizes)
DATASETS = {
    "fiq": ("fiq", 2017, 3817, 640, (10, 50), "configs[0]/[1] FashionIQ dress val shape, RN50x4 640-d"),
    "fiq512-dress": ("fiq", 2017, 3817, 512, (10, 50), "configs[1] FashionIQ dress val, ViT-B-16 512-d"),
    "fiq512-shirt": ("fiq", 2038, 6346, 512, (10, 50), "configs[1] FashionIQ shirt val, ViT-B-16 512-d"),
    "fiq512-toptee": ("fiq", 1961, 5373, 512, (10, 50), "configs[1] FashionIQ toptee val, ViT-B-16 512-d"),
    "shoes": ("fiq", 1761, 4658, 640, (10, 50), "Shoes val shape, 640-d"),
    "f200k": ("200k", 33480, 29789, 640, (1, 10, 50), "configs[2] Fashion200k test shape, 640-d, any-hit recall"),
    "cirr": ("cirr", 4181, 2297, 640, (1, 5, 10, 50), "configs[3] CIRR val shape, 640-d, Recall@K + subset Recall_s@K"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="scale", choices=["scale"] + sorted(DATASETS),
                    help="scale = BASELINE.json configs[4] (default); the others are the dataset shapes of configs[0..3]")
    ap.add_argument("--gallery-rows", type=int, default=100_000_000, help="total gallery rows over all ranks")
    ap.add_argument("--gallery-order", default="random", choices=["random", "clustered"],
                    help="clustered: contiguous clusters of 2048 near-duplicate rows (an ordered catalogue)")
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--dim", type=int, default=640)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-sample-queries", type=int, default=128)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU candidate exchange: fused peer-memory stores (p2p) or NCCL all-gather")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p.get("bf16_tflops_sustained", 1400.0)), "bf16_burst": float(p.get("bf16_tflops", 1590.0)),
                "hbm": float(p.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def profiled_traffic(queries, dim, rows_rank):
    """DRAM traffic of one profiled launch of the dominant kernel, read from the committed ncu summary
    (profiles/sim_topk_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` raw CSV).  Only
    returned when the capture's workload matches this run: same batch and width, and a launch of the size this run's
    schedule actually issues (ops.LAUNCH_MAX_ROWS rows, which needs a shard at least that large); otherwise null -- a
    literal from another configuration would be meaningless."""
    from fashionern_aaai2024_b200 import ops
    path = os.path.join(ROOT, "profiles", "sim_topk_traffic.json")
    if not os.path.exists(path) or ops.LAUNCH_MAX_ROWS is None or rows_rank < 2 * ops.LAUNCH_MAX_ROWS:
        return None
    with open(path) as f:
        entries = json.load(f)
    for e in entries:
        if e["queries"] == queries and e["dim"] == dim and e["launch_rows"] == ops.LAUNCH_MAX_ROWS:
            return e
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline of the scaling config: the reference's torch CPU path (oracle port), bounded sample
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_qps(args, steps=1, warmup=0):
    """fp32 `1 - q @ G.T` + FULL torch.argsort + first-k + id-membership recall (run/test/test_fiq.py:49-60) on the
    host cores, on a bounded sample (sq queries x sn gallery rows).  The scoring pass scales linearly in N and is
    scaled to the benched gallery; the query-side fusion head (models/fusion_model.py:86-94) does not depend on N
    and is timed on its own and added unscaled."""
    from oracle import ern_oracle as orc
    from fashionern_aaai2024_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count())
    sq, sn = args.cpu_sample_queries, min(args.cpu_sample_rows, args.gallery_rows)
    g = torch.Generator().manual_seed(5000)
    gal = torch.nn.functional.normalize(torch.randn(sn, args.dim, generator=g), dim=-1)
    pred = torch.nn.functional.normalize(torch.randn(sq, args.dim, generator=g), dim=-1)
    tgt = torch.randint(0, sn, (sq,), generator=g).numpy()
    cls = torch.arange(sn).numpy()
    sd = syn.combiner_state(7, args.dim)
    img, txt = syn.features(8, sq, args.dim), syn.features(9, sq, args.dim)
    t_head, t_score = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            _ = orc.combiner_forward(sd, img, txt)                     # models/fusion_model.py:86-94 (query fusion)
        t1 = time.perf_counter()
        d = orc.distances(pred, gal)                                   # :49
        order = torch.argsort(d, dim=-1)                               # :50 (full sort, as the reference does)
        ranks = orc.first_hit_rank(order[:, :args.k], cls, tgt)        # :51-55 restated on ids
        _ = orc.recall_at(ranks, KS)                                   # :59-60
        t2 = time.perf_counter()
        if i >= warmup:
            t_head.append(t1 - t0)
            t_score.append(t2 - t1)
    th, ts = sum(t_head) / len(t_head), sum(t_score) / len(t_score)
    t_full = th + ts * args.gallery_rows / sn                          # seconds per sq queries on the benched gallery
    qps = sq / t_full
    return qps, th + ts, {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                          "sample": f"{sq} queries x {sn} gallery rows x {args.dim}-d fp32: fusion head {th:.3f} s (not scaled) + "
                                    f"1 - q@G.T, full argsort, recall {ts:.2f} s (scaled linearly in N to {args.gallery_rows} rows)"}


def workload_config(args, world):
    if args.config != "scale":
        kind, q, n, dim, ks, label = DATASETS[args.config]
        return {"workload": f"{label}: {q} queries x {n} gallery rows, synthetic features, metric tail (gallery "
                            f"normalise + VisualSR + fusion head over the gallery, scoring, top-50, Recall@{list(ks)}"
                            + (", reference removal, subset recall" if kind == "cirr" else "") + ")",
                "dataset": args.config, "queries_per_step": q, "gallery_rows": n, "dim": dim, "k": 50,
                "parallelism": "single GPU",
                "l2": "working set < L2: 512 MB written between timed steps to flush it"}
    return {"workload": f"synthetic gallery {args.gallery_rows} rows x {args.dim}-d bf16 (unit-norm, seeded on device, "
                        f"{args.gallery_order} order), {args.queries}-query batches, top-{args.k}, Recall@{list(KS)}",
            "gallery_rows": args.gallery_rows, "gallery_order": args.gallery_order, "queries_per_step": args.queries,
            "dim": args.dim, "k": args.k,
            "parallelism": f"gallery row-sharded over {world} GPU(s), queries replicated"
                           + (f", candidate exchange: {args.exchange}" if world > 1 else ""),
            "baseline_config": "BASELINE.json configs[4] (synthetic gallery scaling, 1M/10M/100M x 640-d, 4096-query batches, "
                               "top-100); its metric text says top-50 -- the config's harder k = 100 is the default",
            "l2": "gallery shard >> 126 MB L2, streamed from HBM every step (no flush needed)"
                  if args.gallery_rows // world * args.dim * 2 > (1 << 29) else
                  "gallery shard comparable to L2: 512 MB written between timed steps to flush it"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != "scale":
        return run_reference_dataset(args)
    qps, t, cb = cpu_reference_qps(args, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    print(json.dumps({
        "impl": "reference", "metric": f"queries/sec (top-{args.k}, {args.dim}-d)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": args.queries / qps * 1e3, "sample_pass_ms": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus), "cpu_baseline": cb,
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------------------
# dataset configs (BASELINE.json configs[0..3])
# ----------------------------------------------------------------------------------------------------------
def dataset_inputs(name):
    """Seeded synthetic inputs of a dataset config: raw gallery CLIP features, gallery patch features, unit-norm
    predicted query features, names (unique ids, or non-unique captions for Fashion200k), CIRR references/groups."""
    from fashionern_aaai2024_b200 import synthetic as syn
    kind, q, n, dim, ks, _ = DATASETS[name]
    seed = 900 + sorted(DATASETS).index(name) * 10
    g = torch.Generator().manual_seed(seed)
    d = {"kind": kind, "q": q, "n": n, "dim": dim, "ks": ks, "seed": seed,
         "index_features": syn.features(seed + 1, n, dim), "index_local": syn.patch_features(seed + 2, n, dim),
         "pred": syn.features(seed + 3, q, dim, unit=True), "ref_idx": torch.randint(0, n, (q,), generator=g)}
    if kind == "200k":
        d["names"] = syn.caption_names(seed + 6, n, classes=max(8, n // 6))
    elif kind == "cirr":
        d["names"] = syn.unique_names(n, "dev-{}-img")
    else:
        d["names"] = syn.unique_names(n)
    return d


def cirr_groups(seed, ref_idx, tgt_idx, n):
    """6 distinct members per query incl. the reference and the target (dataloader/cirr.py:52,73), shuffled."""
    g = torch.Generator().manual_seed(seed)
    q = len(ref_idx)
    others = torch.randint(0, n, (q, 12), generator=g)
    out = []
    for i in range(q):
        r, t = int(ref_idx[i]), int(tgt_idx[i])
        mem = [r, t]
        for x in others[i].tolist():
            if x not in mem:
                mem.append(x)
            if len(mem) == 6:
                break
        j = 0
        while len(mem) < 6:                      # (astronomically rare) fill deterministically
            if j not in mem:
                mem.append(j)
            j += 1
        perm = torch.randperm(6, generator=g).tolist()
        out.append([mem[p] for p in perm])
    return out


def run_reference_dataset(args):
    """CPU arm of a dataset config: the reference's own `compute_*_val_metrics`, called verbatim where
    /root/reference is mounted (kind "reference": its gallery-side SR_module + Combiner_module, `1 - pred @ G.T`,
    full argsort, numpy string compares; the off-path query encoder is replaced by a table of the same predicted
    features ours gets), else the oracle's restatement of the same tail (kind "port")."""
    from fashionern_aaai2024_b200 import synthetic as syn
    from oracle import ern_oracle as orc
    torch.set_num_threads(os.cpu_count())
    d = dataset_inputs(args.config)
    kind, q, n, dim, ks = d["kind"], d["q"], d["n"], d["dim"], d["ks"]
    names = d["names"]
    sd_sr, sd_cb = syn.visualsr_state(d["seed"] + 20, dim), syn.combiner_state(d["seed"] + 21, dim)
    tgt_idx = torch.randint(0, n, (q,), generator=torch.Generator().manual_seed(d["seed"] + 7))
    if kind == "cirr":
        tgt_idx = torch.where(tgt_idx == d["ref_idx"], (tgt_idx + 1) % n, tgt_idx)
    ref_names = [names[int(i)] for i in d["ref_idx"]]
    tgt_names = [names[int(i)] for i in tgt_idx]
    members = [[names[m] for m in row] for row in cirr_groups(d["seed"] + 8, d["ref_idx"], tgt_idx, n)] if kind == "cirr" else None
    verbatim = os.path.isdir(REFERENCE_ROOT)
    steps = max(1, min(args.steps, 3))
    times = []
    if verbatim:
        from oracle import ref_harness as ref
        ref.install()
        from models.fusion_model import CombinerSimple, VisualSR

        class TableERN(torch.nn.Module):
            """models/model.py:22-75 dispatch; mode="index" is the reference's own modules, mode="test" a table."""
            def __init__(self):
                super().__init__()
                self.SR_module, self.Combiner_module = VisualSR(embed_dim=dim), CombinerSimple(dim, dim * 4, dim * 8)
                self.SR_module.load_state_dict(sd_sr)
                self.Combiner_module.load_state_dict(sd_cb)

            def forward(self, ref_feats=None, ref_local_feats=None, text_feats=None, text_seq_feats=None,
                        tar_feats=None, tar_local_feats=None, mode="train"):
                if mode == "index":
                    return self.Combiner_module(tar_feats, self.SR_module(tar_local_feats))
                return d["pred"][text_feats[:, 0].long()]

        class IndexClip:
            def encode_text(self, tokens, mode="global", visual_emb=None):
                out = torch.zeros(tokens.shape[0], 77 if mode == "seq" else 1, 1)
                out[:, 0, 0] = tokens[:, 0].float()
                return out if mode == "seq" else (out[:, 0], None)

        model = TableERN().eval().float()
        patch = torch.zeros(13, 1)
        ds = ref.FakeRelative("200k" if kind == "200k" else kind, ref_names, tgt_names, [patch] * q, members)
        fn = ref.metric_fn("200k" if kind == "200k" else ("cirr" if kind == "cirr" else "fiq"))
        import contextlib
        for _ in range(steps):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(None), torch.no_grad():
                out = fn(ds, IndexClip(), d["index_features"], d["index_local"], names, model, "cpu", dim, 32, 0, "RN50x4")
            times.append(time.perf_counter() - t0)
    else:
        for _ in range(steps):
            t0 = time.perf_counter()
            with torch.no_grad():
                gal = orc.combiner_forward(sd_cb, orc.gallery_normalize(d["index_features"]),
                                           orc.visual_sr_forward(sd_sr, d["index_local"]))
                if kind == "cirr":
                    out = orc.cirr_metrics(d["pred"], gal, names, ref_names, tgt_names, members)
                elif kind == "200k":
                    out = orc.f200k_metrics(d["pred"], gal, names, tgt_names, (10, 50))
                else:
                    out = orc.fiq_metrics(d["pred"], gal, names, tgt_names, (10, 50))
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    qps = q / t
    cb = {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference" if verbatim else "port",
          "sample": f"whole config ({q} x {n}), {t:.2f} s per pass"
                    + (", verbatim run/test compute_*_val_metrics with the query encoder replaced by a feature table"
                       if verbatim else ", oracle restatement of the tail (no /root/reference on this box)")}
    print(json.dumps({
        "impl": "reference", "metric": f"queries/sec (top-50, {dim}-d)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1), "cpu_baseline": cb, "recall": [float(x) for x in out],
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def flush_l2(buf):
    buf.add_(1)


def run_dataset(args):
    import fashionern_aaai2024_b200 as ern
    from fashionern_aaai2024_b200 import metrics, ops, synthetic as syn
    from fashionern_aaai2024_b200._lib import MODE_BF16, RANK_REFERENCE
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"config": args.config, "unavailable": "dataset configs are single-GPU workloads (< 1 ms of work)"}))
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ern._lib.require_device(dev)
    d = dataset_inputs(args.config)
    kind, q, n, dim, ks = d["kind"], d["q"], d["n"], d["dim"], d["ks"]
    names = d["names"]
    model = ern.ERN(None, dim, None)
    model.load_state_dict(syn.ern_full_state(d["seed"], dim))       # SR_module: seed + 20, Combiner_module: seed + 21
    model = model.to(dev).eval()
    feats_d, local_d = d["index_features"].to(dev), d["index_local"].to(dev)
    pred_h = d["pred"].pin_memory()
    pred_d = pred_h.to(dev)
    cls_np, table, _ = metrics.factorize_names(names)
    cls_d = torch.from_numpy(cls_np).to(dev)
    row_ids = torch.arange(n, dtype=torch.int32, device=dev)
    ref_d = d["ref_idx"].int().to(dev)

    def tail(pred, tgt_cls, tgt_row=None, mem=None):
        gallery = metrics.prepare_gallery(feats_d, local_d, model, dev)                  # test_fiq.py:45-46
        qb, gb, mode = metrics._operands(pred, gallery, "bf16")
        if kind == "cirr":
            vals, ids, _, st = ops.sim_topk(qb, gb, 50, mode=mode, rank_by=RANK_REFERENCE, exclude_ids=ref_d, check_overflow=False)
            counts, _ = ops.recall_at_k(ids, row_ids, tgt_row, (1, 5, 10, 50))
            gcounts, _ = ops.cirr_subset_recall(qb, gb, mem, ref_d, tgt_row, (1, 2, 3), rank_by=RANK_REFERENCE)
            counts = torch.cat([gcounts, counts])
        else:
            vals, ids, _, st = ops.sim_topk(qb, gb, 50, mode=mode, rank_by=RANK_REFERENCE, check_overflow=False)
            counts, _ = ops.recall_at_k(ids, cls_d, tgt_cls, ks)
        return vals, ids, counts, st, qb, gb

    # ---- parity of this run's ranking (same bf16 operands; fp32 matmul + stable sort on the GPU) -------------
    dummy = torch.zeros(q, dtype=torch.int32, device=dev)
    mem0 = torch.zeros((q, 6), dtype=torch.int32, device=dev)
    vals, ids, _, st, qb, gb = tail(pred_d, dummy, dummy, mem0)
    pc = dataset_parity(qb, gb, vals, ids, ref_d if kind == "cirr" else None)
    # ---- plant every query's target at a chosen rank of the verified ranking ---------------------------------
    planted = syn.planted_ranks(11, q, max_rank=50).clamp(max=49).to(dev)
    tgt_row = ids.gather(1, planted[:, None]).squeeze(1).contiguous()
    tgt_cls = cls_d[tgt_row.long()].contiguous()
    mem_d = None
    if kind == "cirr":
        mem_d = torch.tensor(cirr_groups(d["seed"] + 8, d["ref_idx"], tgt_row.cpu(), n), dtype=torch.int32, device=dev)
    tgt_h = (tgt_row if kind == "cirr" else tgt_cls).cpu().pin_memory()
    cnt_h = torch.empty(len(ks) + (3 if kind == "cirr" else 0), dtype=torch.int32).pin_memory()
    scratch = torch.zeros(128 << 20, dtype=torch.float32, device=dev)

    def step_resident():
        return tail(pred_d, tgt_cls, tgt_row, mem_d)

    def step_e2e():
        p = pred_h.to(dev, non_blocking=True)
        t = tgt_h.to(dev, non_blocking=True)
        out = tail(p, t, t, mem_d)
        cnt_h.copy_(out[2], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    def timed(fn, steps):
        ms = []
        for _ in range(steps):
            flush_l2(scratch)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return sum(ms) / len(ms)

    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    t0 = time.time()
    l0 = ops.launch_counter.n
    ms_step = timed(step_resident, args.steps)
    launches = ops.launch_counter.n - l0
    clocks = sampler.stop(t0, time.time())
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    counts = out[2].cpu().tolist()
    status_ok = int(out[3][0].item()) == 0
    recall = [metrics.percent(c, q) for c in counts]
    # expected recall from the planted ranks (unique names: hit iff planted rank < K; any-hit can only be earlier)
    pr = planted.cpu()
    expect = [metrics.percent(int((pr < k).sum()), q) for k in ks]
    rc = recall[3:] if kind == "cirr" else recall
    recall_ok = all((a >= b) if kind == "200k" else (a == b) for a, b in zip(rc, expect))
    pk = peaks()
    flop = 2.0 * q * n * dim + n * (144.0 * dim * dim + 16 * dim) + n * 28.0 * dim * dim
    cpu_baseline = None
    if not args.no_cpu_baseline:
        import io
        import contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            run_reference_dataset(args)
        cpu_baseline = json.loads(buf.getvalue().strip().splitlines()[-1])["cpu_baseline"]
    print(json.dumps({
        "metric": f"queries/sec (top-50, {dim}-d)", "value": q / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, 1),
        "e2e": {"value": q / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": pred_h.numel() * 4 + tgt_h.numel() * 4,
                "d2h_bytes_per_step": cnt_h.numel() * 4, "ms_per_step": ms_e2e,
                "note": "query features + target ids uploaded, hit counts downloaded every step; the gallery's CLIP "
                        "features stay on the device, as in the reference (utils/utils.py:44-69 returns device tensors)"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": flop / (ms_step * 1e-3) / 1e12, "peak": pk["bf16_sustained"],
                     "unit": "TFLOP/s", "frac": flop / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"], "traffic": None,
                     "note": "launch-latency-bound shape (a few GFLOP of scoring, gallery-side heads dominate the FLOP)"},
        "cpu_baseline": cpu_baseline, "recall": recall, "recall_expected": expect, "recall_ok": recall_ok,
        "parity_check": pc, "status_ok": status_ok}))


def dataset_parity(qb, gb, vals, ids, exclude):
    """Whole ranking of a dataset config against fp32 matmul (CUDA cores, TF32 off) + stable sort on the same
    bf16-rounded operands, ranking by the reference's quantity -(1 - s) with ties -> lower id."""
    torch.backends.cuda.matmul.allow_tf32 = False
    q, k = ids.shape
    n = gb.shape[0]
    gf = gb.float()
    max_gap, exact, total, max_err = 0.0, 0, 0, 0.0
    for s in range(0, q, 2048):
        sc = -(1.0 - qb[s:s + 2048].float() @ gf.T)
        if exclude is not None:
            sc.scatter_(1, exclude[s:s + 2048].long()[:, None], float("-inf"))
        so, io = torch.sort(sc, dim=1, descending=True, stable=True)
        so, io = so[:, :k], io[:, :k]
        mine = ids[s:s + 2048].long()
        got = sc.gather(1, mine.clamp(min=0))
        max_gap = max(max_gap, float((got - so).abs().max()))
        max_err = max(max_err, float((got - vals[s:s + 2048]).abs().max()))
        exact += int((mine == io).sum())
        total += mine.numel()
        srt = torch.sort(mine, dim=1).values
        if bool((srt[:, 1:] == srt[:, :-1]).any()) or bool((mine < 0).any()) or bool((mine >= n).any()):
            return {"ok": False, "error": "duplicate or out-of-range id"}
    tol = TOL + 1.2e-7          # ranking by -(1 - s): one more fp32 rounding than s itself
    return {"queries": q, "rows": n, "reference": "torch fp32 matmul (TF32 off) + stable sort of -(1 - s) on the same "
            "bf16-rounded operands, on the GPU, outside the timed region", "tol": tol, "max_rank_gap": max_gap,
            "max_score_err": max_err, "exact_frac": exact / max(total, 1), "ok": max_gap <= tol and max_err <= tol}


# ----------------------------------------------------------------------------------------------------------
# scaling config: gallery generation and the in-run parity check
# ----------------------------------------------------------------------------------------------------------
CLUSTER_ROWS = 2048


def make_gallery(rows, dim, dev, seed, order):
    """Unit-norm bf16 gallery shard generated on the device in row blocks.
    `clustered`: contiguous clusters of 2048 rows = normalize(centre + 0.6 * unit noise) (pairwise cosine 0.74 inside
    a cluster) -- a catalogue stored product by product, so the survivors of a query arrive in bursts."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    gallery = torch.empty((rows, dim), dtype=torch.bfloat16, device=dev)
    blk = 1 << 20
    for s in range(0, rows, blk):
        nb = min(blk, rows - s)
        x = torch.nn.functional.normalize(torch.randn(nb, dim, generator=gen, device=dev), dim=-1)
        if order == "clustered":
            ncl = (nb + CLUSTER_ROWS - 1) // CLUSTER_ROWS
            c = torch.nn.functional.normalize(torch.randn(ncl, dim, generator=gen, device=dev), dim=-1)
            x = torch.nn.functional.normalize(c.repeat_interleave(CLUSTER_ROWS, 0)[:nb] + 0.6 * x, dim=-1)
        gallery[s:s + nb] = x.bfloat16()
    return gallery


def parity_check(qb, gallery, begin, vals, ids, world, dev, dist):
    """Is (vals, ids) -- the global top-k every rank holds -- the exact top-k of the sharded gallery?

    Every rank recomputes ALL scores of its shard with a cuBLAS bf16 -> fp32 GEMM on the same bf16 operands and counts,
    per query, the rows above the claimed k-th value; the counts are summed over the ranks.  With v_k the claimed k-th
    value:  (a) every claimed id is found in exactly one shard and its recomputed score matches the claimed value,
    (b) the number of rows scoring > v_k + tol equals the number of claimed ids scoring > v_k + tol (nothing that
    belongs in the list is missing), (c) at least k rows score >= v_k - tol, (d) the list is sorted by value
    descending / id ascending without duplicates.  Then 16 queries are re-ranked with an fp32 CUDA-core matmul +
    torch.topk over the shard(s) and compared position by position (gap-aware)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    Q, K = ids.shape
    rows = gallery.shape[0]
    vk = vals[:, K - 1].contiguous()
    hi, lo = (vk + TOL)[:, None], (vk - TOL)[:, None]
    cnt_gt = torch.zeros(Q, dtype=torch.int64, device=dev)
    cnt_ge = torch.zeros(Q, dtype=torch.int64, device=dev)
    found = torch.zeros(Q, dtype=torch.int64, device=dev)
    claimed_gt = torch.zeros(Q, dtype=torch.int64, device=dev)
    max_err = torch.zeros(1, device=dev)
    ids64 = ids.long()
    NS = 16
    sub_vals = torch.full((NS, K), float("-inf"), device=dev)
    sub_ids = torch.full((NS, K), -1, dtype=torch.int64, device=dev)
    q_sub = qb[:NS].float()
    chunk = 1 << 18
    for s in range(0, rows, chunk):
        g = gallery[s:s + chunk]
        c = g.shape[0]
        sc = torch.mm(qb, g.T, out_dtype=torch.float32)                     # [Q, c] every score of this block
        cnt_gt += (sc > hi).sum(1)
        cnt_ge += (sc >= lo).sum(1)
        loc = ids64 - (begin + s)
        mine = (loc >= 0) & (loc < c)
        got = sc.gather(1, loc.clamp(0, c - 1))
        found += mine.sum(1)
        claimed_gt += (mine & (got > hi)).sum(1)
        max_err = torch.maximum(max_err, torch.where(mine, (got - vals).abs(), torch.zeros_like(got)).max())
        # 16-query subset: fp32 CUDA-core matmul (no tensor cores) + running top-k
        s32 = q_sub @ g.float().T
        kk = min(K, c)
        tv, ti = torch.topk(s32, kk, dim=1)
        cat_v = torch.cat([sub_vals, tv], 1)
        cat_i = torch.cat([sub_ids, ti + begin + s], 1)
        order = torch.sort(cat_v, dim=1, descending=True, stable=True).indices[:, :K]
        sub_vals, sub_ids = cat_v.gather(1, order), cat_i.gather(1, order)
        del sc, s32
    if world > 1:
        for t in (cnt_gt, cnt_ge, found, claimed_gt):
            dist.all_reduce(t)
        dist.all_reduce(max_err, op=dist.ReduceOp.MAX)
        gv = [torch.empty_like(sub_vals) for _ in range(world)]
        gi = [torch.empty_like(sub_ids) for _ in range(world)]
        dist.all_gather(gv, sub_vals)
        dist.all_gather(gi, sub_ids)
        cat_v, cat_i = torch.cat(gv, 1), torch.cat(gi, 1)
        order = torch.sort(cat_v, dim=1, descending=True, stable=True).indices[:, :K]
        sub_vals, sub_ids = cat_v.gather(1, order), cat_i.gather(1, order)
    srt = torch.sort(ids64, dim=1).values
    dup = bool((srt[:, 1:] == srt[:, :-1]).any())
    v0, v1, i0, i1 = vals[:, :-1], vals[:, 1:], ids64[:, :-1], ids64[:, 1:]
    ordered = bool(((v0 > v1) | ((v0 == v1) & (i0 < i1))).all())
    missed = int((cnt_gt - claimed_gt).abs().sum())
    not_found = int((found != K).sum())
    short = int((cnt_ge < K).sum())
    # subset: the value we report at rank j must be within tol of the fp32 reference's rank-j value (gap-aware)
    rank_gap = float((vals[:NS] - sub_vals).abs().max())
    same = float((ids64[:NS] == sub_ids).float().mean())
    err = float(max_err)
    ok = (not dup) and ordered and missed == 0 and not_found == 0 and short == 0 and err <= TOL and rank_gap <= TOL
    return {"queries": Q, "rows_per_rank": rows, "ranks": world, "tol": TOL,
            "reference": "all scores of every shard recomputed with cuBLAS bf16->fp32 (torch.mm out_dtype=float32) on the "
                         "same bf16 operands; rows above the claimed k-th value counted and summed over ranks; "
                         f"{NS} queries re-ranked with fp32 CUDA-core matmul + torch.topk (+ all-gather/merge over ranks)",
            "missed_rows": missed, "claimed_ids_not_found": not_found, "queries_short_of_k": short,
            "near_ties_at_kth": int((cnt_ge - K).clamp(min=0).sum()), "duplicates": dup, "sorted": ordered,
            "max_score_err": err, "subset_queries": NS, "max_rank_gap": rank_gap, "subset_exact_frac": same, "ok": ok}


# ----------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "scale":
        return run_dataset(args)

    import torch.distributed as dist
    import fashionern_aaai2024_b200 as ern
    from fashionern_aaai2024_b200 import ops, sharded, synthetic as syn
    from fashionern_aaai2024_b200.combiner import CombinerSimple

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ern._lib.require_device(dev)

    Q, D, K, N = args.queries, args.dim, args.k, args.gallery_rows
    begin, end = sharded.shard_bounds(N, world, rank)
    rows = end - begin
    gallery = make_gallery(rows, D, dev, 5000 + rank, args.gallery_order)
    class_of = torch.arange(N, dtype=torch.int32, device=dev)          # unique names: class id == global row id
    small = rows * D * 2 <= (1 << 29)
    scratch = torch.zeros(128 << 20, dtype=torch.float32, device=dev) if small else None

    # ---- query batch (host, pinned): reference-image + text CLIP features and target ids ---------------------
    head = CombinerSimple(D, 4 * D, 8 * D, mode="bf16")
    head.load_state_dict(syn.combiner_state(7, D))
    head = head.to(dev).eval()
    img_h = syn.features(8, Q, D).pin_memory()
    txt_h = syn.features(9, Q, D).pin_memory()
    tgt_h = torch.randint(0, N, (Q,), generator=torch.Generator().manual_seed(10), dtype=torch.int32).pin_memory()
    img_d, txt_d, tgt_d = img_h.to(dev), txt_h.to(dev), tgt_h.to(dev)
    ids_h = torch.empty((Q, K), dtype=torch.int32).pin_memory()
    val_h = torch.empty((Q, K), dtype=torch.float32).pin_memory()
    cnt_h = torch.empty(len(KS), dtype=torch.int32).pin_memory()
    statuses = []

    def step(img, txt, tgt, timed_sim=None):
        with torch.no_grad():
            _, qb = head(img, txt, want_bf16=True)                       # fusion head -> bf16 unit-norm queries
        if timed_sim is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        vals, ids, _keys, status = sharded.sharded_topk(qb, gallery, K, begin, check_overflow=False, exchange=args.exchange)
        if timed_sim is not None:
            e1.record()
            timed_sim.append((e0, e1))
        counts, _ranks = ops.recall_at_k(ids, class_of, tgt, KS)
        statuses.append(status)
        return vals, ids, counts, status, qb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between two barriers; large shards: one event pair around all K steps; shards that could sit in
        L2: the L2 is flushed before every step and the K per-step event intervals are summed."""
        barrier()
        pairs = []
        t0 = time.time()
        for _ in range(steps):
            if scratch is not None:
                flush_l2(scratch)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            pairs.append((e0, e1))
        barrier()
        ms = pairs[0][0].elapsed_time(pairs[-1][1]) if scratch is None else sum(a.elapsed_time(b) for a, b in pairs)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, time.time()

    # ---- first pass: verify its top-k against an independent recomputation, then plant the targets from it ------
    out = step(img_d, txt_d, tgt_d)
    pc = None
    if not args.no_parity_check:
        pc = parity_check(out[4], gallery, begin, out[0], out[1], world, dev, dist)
        torch.cuda.empty_cache()
    planted = syn.planted_ranks(11, Q, max_rank=K).clamp(max=K - 1).to(dev)
    tgt_d = out[1].gather(1, planted[:, None]).squeeze(1).contiguous()
    tgt_h.copy_(tgt_d)
    for _ in range(max(args.warmup, 3)):
        out = step(img_d, txt_d, tgt_d)
    torch.cuda.synchronize()
    statuses.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = ops.launch_counter.n
    events = []
    ms_total, t0, t1 = timed(lambda: step(img_d, txt_d, tgt_d, events), args.steps)
    launches = ops.launch_counter.n - l0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    sim_ms = [a.elapsed_time(b) for a, b in events]
    ms_step = ms_total / args.steps
    value = Q / (ms_step * 1e-3)

    # ---- `e2e`: host inputs, pinned H2D per step, results read back per step ---------------------------------
    # A serving loop: the H2D copy of batch i+1 is issued on a copy stream while batch i computes (double-buffered
    # device inputs) and the host consumes the downloaded result of batch i-1 while batch i runs (double-buffered
    # pinned outputs).  Every step uploads its own inputs and downloads its own results, and all K uploads and
    # downloads (including the host wait for the last result) happen inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    dev_in = [(torch.empty_like(img_d), torch.empty_like(txt_d), torch.empty_like(tgt_d)) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def issue_upload(i):
        slot = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[slot])                        # the batch that used this slot has finished
            for dst, src in zip(dev_in[slot], (img_h, txt_h, tgt_h)):
                dst.copy_(src, non_blocking=True)
            ev_ready[slot].record(copy_stream)

    out_h = [(ids_h, val_h, cnt_h), (torch.empty_like(ids_h).pin_memory(), torch.empty_like(val_h).pin_memory(),
                                     torch.empty_like(cnt_h).pin_memory())]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(steps):
        cur = torch.cuda.current_stream()
        for e in ev_free:
            e.record(cur)
        issue_upload(0)
        for i in range(steps):
            slot = i & 1
            cur.wait_event(ev_ready[slot])
            if i + 1 < steps:
                issue_upload(i + 1)
            img, txt, tgt = dev_in[slot]
            if scratch is not None:
                flush_l2(scratch)                                        # (inside the e2e interval: counts against us)
            vals, ids, counts, _, _ = step(img, txt, tgt)
            ev_free[slot].record(cur)
            for dst, src in zip(out_h[slot], (ids, vals, counts)):       # this batch's result -> pinned host memory
                dst.copy_(src, non_blocking=True)
            ev_out[slot].record(cur)
            if i > 0:
                ev_out[slot ^ 1].synchronize()                           # the host consumes batch i-1 while batch i runs
        ev_out[(steps - 1) & 1].synchronize()                            # ... and the last one

    e2e_loop(2)
    ms_e2e, _, _ = timed(lambda: e2e_loop(args.steps), 1)
    e2e_value = Q / (ms_e2e / args.steps * 1e-3)
    h2d = img_h.numel() * 4 + txt_h.numel() * 4 + tgt_h.numel() * 4
    d2h = ids_h.numel() * 4 + val_h.numel() * 4 + cnt_h.numel() * 4
    # the candidate store of EVERY timed / e2e step on EVERY rank must have stayed consistent
    bad = torch.stack([s[0] for s in statuses]).ne(0).sum().reshape(1).to(torch.int64)
    if world > 1:
        dist.all_reduce(bad)
    status_ok = int(bad.item()) == 0

    # ---- roofline of the dominant kernel (sim_topk_tc_kernel, all launches of one step on this rank) ---------
    pk = peaks()
    flop_rank = 2.0 * Q * rows * D                                        # 2*D FLOP per (query,row)
    sim_avg_ms = sum(sim_ms) / max(len(sim_ms), 1)
    achieved = flop_rank / (sim_avg_ms * 1e-3) / 1e12
    tr = profiled_traffic(Q, D, rows)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"],
                "traffic": (tr["dram_bytes_read"] + tr["dram_bytes_write"]) if tr else None,
                "traffic_algorithmic": tr["algorithmic_bytes"] if tr else None,
                "traffic_basis": (f"one launch of {tr['launch_rows']} gallery rows x {tr['queries']} queries (the launch size of this run's schedule), ncu --set full "
                                  f"({tr['source']}), read by bench.py from profiles/sim_topk_traffic.json") if tr else
                                 "no committed ncu capture matches this configuration",
                "achieved_basis": "2*Q*rows*D FLOP of this rank's shard / CUDA-event time of the scoring call of one step "
                                  "(all launches of the kernel plus the interleaved selection launches)",
                "kernel": "ern::simtc::sim_topk_tc_kernel (all launches of one step incl. the interleaved "
                          "select_topk_kernel launches, CUDA events on the launching stream)",
                "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "frac_of_burst_peak": achieved / pk["bf16_burst"],
                "hbm_gbs_achieved": rows * D * 2 / (sim_avg_ms * 1e-3) / 1e9, "hbm_peak_gbs": pk["hbm"]}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            _, _, cpu_baseline = cpu_reference_qps(args)
        recall = [float(100.0 * c / Q) for c in out[2].cpu().tolist()]
        pr = planted.cpu()
        expect = [float(100.0 * int((pr < k).sum()) / Q) for k in KS]
        if world > 1 and sharded.exchange_in_use(args.exchange) != args.exchange:
            args.exchange = sharded.exchange_in_use(args.exchange) + " (p2p unavailable)"
        line = {
            "metric": f"queries/sec (top-{K}, {D}-d)", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_check": pc,
            "recall_at": dict(zip([str(k) for k in KS], recall)),
            "recall_expected": dict(zip([str(k) for k in KS], expect)),
            "recall_basis": "targets planted at seeded ranks of the ranking verified by parity_check; expected = share "
                            "of planted ranks < K",
            "status_ok": status_ok,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
