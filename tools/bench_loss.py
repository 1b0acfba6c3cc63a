"""Device-side timing of the training criterion: forward + backward of the in-batch classification
loss at the reference's batch size (run/train/train_fiq.py:194 --batch-size 1024) and larger, beside the reference's
own torch formulation (losses/loss.py:10-14 under fp16 autocast + autograd) on the same GPU."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern  # noqa: E402
from fashionern_aaai2024_b200 import synthetic as syn  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    out = []
    for dim in (640, 512):
        for rows in (256, 1024, 4096, 16384):
            pred, tar = syn.loss_pair(3, rows, dim)
            p = pred.to(dev).requires_grad_(True)
            t = tar.to(dev).requires_grad_(True)
            labels = torch.arange(rows, device=dev)

            def torch_step():
                p.grad = t.grad = None
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = F.cross_entropy(100 * p @ t.T, labels)
                (loss * 1024.0).backward()
                return loss

            rec = {"op": "BatchBasedClassificationLoss fwd+bwd", "dim": dim, "rows": rows,
                   "torch_autocast_ms": timeit(torch_step)}
            for prec in ("bf16", "fp32"):
                if prec == "fp32" and rows > 4096:
                    continue
                crit = ern.BatchBasedClassificationLoss(precision=prec)

                def step():
                    p.grad = t.grad = None
                    loss = crit(p, t)
                    (loss * 1024.0).backward()
                    return loss

                rec[f"{prec}_ms"] = timeit(step)
                rec[f"{prec}_loss"] = float(step())
            rec["torch_loss"] = float(torch_step())
            # 3 logits GEMMs (forward + 2 recomputations) + 2 gradient GEMMs, 2*B*B*D each
            rec["bf16_tflops"] = 5 * 2.0 * rows * rows * dim / rec["bf16_ms"] / 1e9
            out.append(rec)
            print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_loss.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
