"""One forward + backward of the training criterion per batch size, for ncu launch lists (tools/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern  # noqa: E402
from fashionern_aaai2024_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
for rows in [int(a) for a in sys.argv[1:]] or [1024]:
    pred, tar = syn.loss_pair(3, rows, 640)
    p = pred.to(dev).requires_grad_(True)
    t = tar.to(dev).requires_grad_(True)
    crit = ern.BatchBasedClassificationLoss()
    for _ in range(2):
        p.grad = t.grad = None
        loss = crit(p, t)
        loss.backward()
    torch.cuda.synchronize()
    print(rows, float(loss.detach()))
