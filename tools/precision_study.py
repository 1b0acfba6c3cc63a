"""bf16 product path vs fp32 validation path on UN-rounded features (SURVEY.md P3 / BASELINE.md section 4):
how much does bf16 operand rounding move the top-100 at 1M gallery rows?"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fashionern_aaai2024_b200 import ops  # noqa: E402
from fashionern_aaai2024_b200._lib import MODE_BF16, MODE_FP32  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    q, n, dim, k = 256, 1_000_000, 640, 100
    gen = torch.Generator(device=dev).manual_seed(3)
    gal = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen, device=dev), dim=-1)
    pred = torch.nn.functional.normalize(torch.randn(q, dim, generator=gen, device=dev), dim=-1)
    v32, i32, _, _ = ops.sim_topk(pred, gal, k, mode=MODE_FP32)
    vb, ib, _, _ = ops.sim_topk(pred.bfloat16(), gal.bfloat16(), k, mode=MODE_BF16)
    same_pos = float((i32 == ib).float().mean())
    overlap = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i32.cpu(), ib.cpu())) / (q * k)
    top1 = float((i32[:, 0] == ib[:, 0]).float().mean())
    top10 = sum(len(set(a[:10].tolist()) & set(b[:10].tolist())) for a, b in zip(i32.cpu(), ib.cpu())) / (q * 10)
    rel = float(((vb - v32).abs() / v32.abs()).max())
    gap = float((v32[:, :-1] - v32[:, 1:]).median())
    print(json.dumps({"queries": q, "gallery_rows": n, "dim": dim, "k": k, "same_position_frac": same_pos,
                      "set_overlap_frac": overlap, "top1_agree_frac": top1, "top10_set_overlap_frac": top10,
                      "max_rel_score_diff_rankwise": rel, "median_neighbour_gap_fp32": gap}))


if __name__ == "__main__":
    main()
