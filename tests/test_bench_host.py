"""CPU: the host-side pieces of bench.py -- workload construction, the roofline.traffic lookup, and the reference arm
of a dataset config (the reference's own compute_fiq_val_metrics called verbatim where /root/reference is mounted, the
oracle port elsewhere)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_cirr_groups_hold_reference_and_target_once():
    g = torch.Generator().manual_seed(1)
    n, q = 50, 200
    ref = torch.randint(0, n, (q,), generator=g)
    tgt = (ref + 1 + torch.randint(0, n - 1, (q,), generator=g)) % n
    groups = bench.cirr_groups(3, ref, tgt, n)
    for r, t, mem in zip(ref.tolist(), tgt.tolist(), groups):
        assert len(mem) == 6 and len(set(mem)) == 6 and r in mem and t in mem and all(0 <= m < n for m in mem)


def test_dataset_inputs_follow_baseline_shapes():
    d = bench.dataset_inputs("cirr")
    assert (d["q"], d["n"], d["dim"]) == (4181, 2297, 640) and d["index_local"].shape == (2297, 13, 640)
    assert len(set(d["names"])) == 2297                                   # CIRR names are unique
    f = bench.dataset_inputs("fiq512-shirt")
    assert (f["q"], f["n"], f["dim"]) == (2038, 6346, 512)
    cfg = bench.workload_config(type("A", (), {"config": "f200k"})(), 1)
    assert cfg["queries_per_step"] == 33480 and cfg["gallery_rows"] == 29789


def test_traffic_is_only_reported_for_the_profiled_configuration():
    from fashionern_aaai2024_b200 import ops
    cap = ops.LAUNCH_MAX_ROWS
    hit = bench.profiled_traffic(4096, 640, 100_000_000)
    assert hit is not None and hit["launch_rows"] == cap and hit["dram_bytes_read"] > hit["algorithmic_bytes"] * 0.9
    assert os.path.exists(os.path.join(ROOT, hit["source"]))               # the ncu CSV it came from is committed
    assert bench.profiled_traffic(512, 640, 100_000_000) is None           # another batch size
    assert bench.profiled_traffic(4096, 512, 100_000_000) is None          # another width
    assert bench.profiled_traffic(4096, 640, 1_000_000) is None            # a shard smaller than the profiled launch


def test_reference_arm_of_a_dataset_config_runs_on_the_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "fiq",
                          "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["value"] > 0
    want_kind = "reference" if os.path.isdir(bench.REFERENCE_ROOT) else "port"
    assert line["cpu_baseline"]["kind"] == want_kind
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and len(line["recall"]) == 2
