"""What fraction of top-k positions is IDENTICAL to the oracle's in the parametrised cases of tests/test_gpu_sim.py?
(The tests' real guard is oracle.compare_topk's rank-wise 2e-6 gap rule; this census is what their secondary
`exact_frac` thresholds are set from.)  Prints one JSON line per case and the minimum per group."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_sim as T  # noqa: E402
from fashionern_aaai2024_b200._lib import MODE_BF16, MODE_FP32  # noqa: E402


def cases_of(fn):
    for m in fn.pytestmark:
        if m.name == "parametrize":
            return list(m.args[1])
    return []


def main():
    dev = torch.device("cuda", 0)
    groups = {"fp32": (T.test_fp32_validation_mode_matches_oracle, MODE_FP32, torch.float32),
              "bf16_single": (T.test_bf16_single_cta_matches_oracle, MODE_BF16, torch.bfloat16),
              "bf16_pair": (T.test_bf16_cta_pair_matches_oracle, MODE_BF16, torch.bfloat16),
              "fp16": (T.test_fp16_operands_match_oracle, MODE_BF16, torch.float16)}
    for name, (fn, mode, cast) in groups.items():
        worst = 1.0
        for q, n, dim, k in cases_of(fn):
            stats, *_ = T.run_case(dev, q, n, dim, k, mode, cast=cast)
            worst = min(worst, stats["exact_frac"])
            print(json.dumps({"group": name, "q": q, "n": n, "dim": dim, "k": k, **stats}))
        print(json.dumps({"group": name, "min_exact_frac": worst}))


if __name__ == "__main__":
    main()
