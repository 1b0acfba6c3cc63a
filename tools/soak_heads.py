"""One-off randomized soak of the secondary kernels against the CPU oracle (not part of the test suite):
fusion head, VisualSR, DVR encoder and the training criterion at random batch sizes, both arithmetic modes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fashionern_aaai2024_b200 as ern  # noqa: E402
from fashionern_aaai2024_b200 import ops, synthetic as syn  # noqa: E402
from fashionern_aaai2024_b200._lib import MODE_BF16, MODE_FP32  # noqa: E402
from oracle import ern_oracle as orc  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", 0)
    worst = {}

    failures = []

    def note(name, err, tol):
        worst[name] = max(worst.get(name, 0.0), err / tol)
        if err > tol:                                  # keep going: the summary lists every case over its tolerance
            failures.append((name, err, tol))

    for it in range(24):
        dim = int(rng.choice([512, 640]))
        mode = "bf16" if it % 3 else "fp32"
        tol = 1e-2 if mode == "bf16" else 1e-5
        rows = int(rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 300, 777, 1000, 2049]))
        sd = syn.combiner_state(100 + it, dim)
        head = ern.CombinerSimple(dim, 4 * dim, 8 * dim, mode=mode)
        head.load_state_dict(sd)
        head = head.to(dev).eval()
        a, b = syn.features(200 + it, rows, dim), syn.features(300 + it, rows, dim)
        with torch.no_grad():
            out = head(a.to(dev), b.to(dev)).cpu()
        note(f"combiner/{mode}", float((out - orc.combiner_forward(sd, a, b)).norm(dim=-1).max()), tol)

        srd = syn.visualsr_state(400 + it, dim)
        sr = ern.VisualSR(dim, mode=mode)
        sr.load_state_dict(srd)
        sr = sr.to(dev).eval()
        x = syn.patch_features(500 + it, rows, dim)
        with torch.no_grad():
            o2 = sr(x.to(dev)).cpu()
        note(f"visualsr/{mode}", float((o2 - orc.visual_sr_forward(srd, x)).norm(dim=-1).max()), tol)

        lm = MODE_BF16 if mode == "bf16" else MODE_FP32
        pred, tar = syn.loss_pair(600 + it, rows, dim)
        rp, rt = (pred.bfloat16().float(), tar.bfloat16().float()) if mode == "bf16" else (pred, tar)
        ref = orc.bbc_loss(rp, rt)
        loss, lse = ops.bbc_loss_forward(pred.to(dev), tar.to(dev), 100.0, lm)
        dp, dt = ops.bbc_loss_backward(pred.to(dev), tar.to(dev), lse, None, 100.0, lm)
        note(f"loss/{mode}", abs(float(loss) - ref[0]), 2e-5 * abs(ref[0]) + 1e-5)
        gmax = max(float(np.abs(ref[2]).max()), float(np.abs(ref[3]).max()))
        floor = (100.0 / rows) * 1.6e-5 * float(tar.abs().max())
        gtol = (1e-2 if mode == "bf16" else 3e-4) * gmax + floor
        note(f"loss-grad/{mode}", max(float(np.abs(dp.cpu().numpy() - ref[2]).max()),
                                      float(np.abs(dt.cpu().numpy() - ref[3]).max())), gtol)

    for it in range(8):
        dim = int(rng.choice([512, 640]))
        mode = "bf16" if it % 2 else "fp32"
        rows = int(rng.choice([1, 3, 17, 32, 65, 130]))
        sd = syn.dvr_full_state(700 + it, dim)
        m = ern.DVR_module(dim, mode=mode)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        m.max_batch = int(rng.choice([16, 64, 2048]))
        patches, tokens = syn.patch_features(800 + it, rows, dim), syn.token_features(900 + it, rows, dim)
        rg, tg = syn.features(1000 + it, rows, dim), syn.features(1100 + it, rows, dim)
        with torch.no_grad():
            out = m(patches.to(dev), tokens.to(dev), rg.to(dev), tg.to(dev)).cpu()
        ref = orc.dvr_forward(sd, patches, tokens, rg, tg)
        note(f"dvr/{mode}", float((out - ref).norm(dim=-1).max()), 1e-2 if mode == "bf16" else 2e-5)
    print("SOAK_OK" if not failures else "SOAK_OVER_TOLERANCE", {k: round(v, 3) for k, v in sorted(worst.items())},
          "(worst error as a fraction of its tolerance)", failures)


if __name__ == "__main__":
    main()
