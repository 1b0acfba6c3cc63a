"""GPU: bench.py's in-run parity check must be able to FAIL.

The headline number verifies its own top-k before timing (bench.parity_check: every score of the shard recomputed with a
cuBLAS bf16->fp32 GEMM, rows above the claimed k-th value counted, 16 queries re-ranked in fp32).  A check that cannot
fail proves nothing, so this test feeds it correct results (ok) and then results with one realistic defect at a time --
a dropped candidate (what a broken filter / merge produces), a wrong score, a swapped order, a duplicated id -- and
requires `ok: false` with the right counter raised.  Also runs the bench command end to end on a small gallery."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(cuda_device):
    import bench
    from fashionern_aaai2024_b200 import ops
    q, n, d, k = 300, 70_000, 128, 20
    gal = bench.make_gallery(n, d, cuda_device, 11, "clustered")
    gen = torch.Generator(device=cuda_device).manual_seed(12)
    qb = torch.nn.functional.normalize(torch.randn(q, d, generator=gen, device=cuda_device), dim=-1).bfloat16()
    vals, ids, _, status = ops.sim_topk(qb, gal, k)
    assert status.cpu().tolist() == [0, 0, 0, 0]
    return bench, qb, gal, vals, ids


def check(case, vals, ids):
    bench, qb, gal, _, _ = case
    return bench.parity_check(qb, gal, 0, vals, ids, 1, qb.device, None)


def test_correct_result_passes(case):
    pc = check(case, case[3], case[4])
    assert pc["ok"] and pc["missed_rows"] == 0 and pc["claimed_ids_not_found"] == 0 and pc["max_score_err"] <= 2e-6
    assert pc["max_rank_gap"] <= 2e-6 and pc["sorted"] and not pc["duplicates"]


def test_dropped_candidate_is_caught(case):
    # the 3rd best of query 5 is lost and everything below moves up one place (a filter that dropped a survivor)
    vals, ids = case[3].clone(), case[4].clone()
    vals[5, 2:-1], ids[5, 2:-1] = case[3][5, 3:], case[4][5, 3:]
    vals[5, -1], ids[5, -1] = case[3][5, -1] - 0.01, (case[4][5, -1] + 1) % 70_000
    pc = check(case, vals, ids)
    assert not pc["ok"] and (pc["missed_rows"] >= 1 or pc["max_score_err"] > 2e-6)


def test_wrong_score_is_caught(case):
    vals = case[3].clone()
    vals[7, 4] += 1e-4
    pc = check(case, vals, case[4])
    assert not pc["ok"] and pc["max_score_err"] > 2e-6


def test_swapped_order_is_caught(case):
    vals, ids = case[3].clone(), case[4].clone()
    vals[9, [0, 1]], ids[9, [0, 1]] = case[3][9, [1, 0]], case[4][9, [1, 0]]
    pc = check(case, vals, ids)
    assert not pc["ok"] and not pc["sorted"]


def test_duplicate_id_is_caught(case):
    ids = case[4].clone()
    ids[3, 6] = ids[3, 5]
    pc = check(case, case[3], ids)
    assert not pc["ok"] and pc["duplicates"]


def test_bench_command_small_gallery():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gallery-rows", "300000", "--queries", "512",
                          "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--gallery-order", "clustered"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["parity_check"]["ok"] and line["status_ok"] and line["gpu_launches"] > 0
    assert line["recall_at"] == line["recall_expected"]
    assert {"roofline", "e2e", "clocks", "config"} <= set(line)
