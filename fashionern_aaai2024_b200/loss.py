"""Drop-in for the reference's training criterion.

``BatchBasedClassificationLoss`` keeps the reference's class name, constructor and ``forward(predicted_features,
tar_features)`` signature (losses/loss.py:6-14) and returns a 0-d tensor that takes part in autograd, so
``self.scaler.scale(loss).backward()`` (run/train/train_fiq.py:134-137) works unchanged.  Forward and backward both
run in libern_b200.so (tcgen05 GEMMs with logsumexp / softmax-gradient epilogues); there is no torch fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import MODE_BF16, MODE_FP32, ErnError


class _BbcLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tar, scale, mode):
        p = pred.detach().float()
        t = tar.detach().float()
        loss, lse = ops.bbc_loss_forward(p, t, scale, mode)
        ctx.save_for_backward(p, t, lse)
        ctx.scale, ctx.mode = scale, mode
        ctx.in_dtypes = (pred.dtype, tar.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        p, t, lse = ctx.saved_tensors
        dpred, dtar = ops.bbc_loss_backward(p, t, lse, grad_out, ctx.scale, ctx.mode)
        return dpred.to(ctx.in_dtypes[0]), dtar.to(ctx.in_dtypes[1]), None, None


class BatchBasedClassificationLoss(nn.Module):
    """``F.cross_entropy(100 * predicted_features @ tar_features.T, arange(B))`` (losses/loss.py:10-14).

    ``precision="bf16"`` (default) rounds the operands to bf16 and accumulates in fp32 on the tensor cores -- the
    counterpart of the fp16 autocast region the reference computes this loss in (run/train/train_fiq.py:124-134),
    except that the logits stay in fp32;  ``precision="fp32"`` is the validation path."""

    def __init__(self, precision: str = "bf16", scale: float = 100.0):
        super().__init__()
        if precision not in ("bf16", "fp32"):
            raise ErnError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.mode = MODE_BF16 if precision == "bf16" else MODE_FP32
        self.scale = float(scale)

    def forward(self, predicted_features: torch.Tensor, tar_features: torch.Tensor) -> torch.Tensor:
        if predicted_features.dim() != 2 or predicted_features.shape != tar_features.shape:
            raise ErnError("predicted_features / tar_features must both be [B, D]")
        if not predicted_features.is_cuda:
            raise ErnError("the B200 loss runs on CUDA tensors only (no CPU fallback)")
        return _BbcLossFn.apply(predicted_features, tar_features, self.scale, self.mode)
