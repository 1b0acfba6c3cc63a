"""GPU: the training criterion -- forward and gradient -- against the reference golden
(losses/loss.py + torch autograd, run on CPU by oracle/make_golden.py) and the float64 oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ern_oracle as orc
from fashionern_aaai2024_b200 import BatchBasedClassificationLoss, ErnError, ops
from fashionern_aaai2024_b200 import synthetic as syn
from fashionern_aaai2024_b200._lib import MODE_BF16, MODE_FP32
from helpers import load_golden

pytestmark = pytest.mark.gpu

# fp32 validation mode: relative error of the loss; error of the gradients relative to their largest entry.  The
# logits are ~100 (ulp 7.6e-6 in fp32), so any fp32 evaluation -- the reference's included: its golden differs from
# the float64 oracle by 1e-4 of the largest gradient entry (tests/golden/pin_report.json) -- carries that much noise
# in the softmax; 3e-4 is that floor with margin.
TOL_F32 = 1e-5
TOL_F32_GRAD = 3e-4
# bf16 mode, against the oracle evaluated on the SAME bf16-rounded operands: the logits are exact fp32 accumulations,
# so the loss agrees to fp32 accuracy; the gradient GEMMs additionally round d(logits) to bf16 (2^-9 per entry)
TOL_BF16_LOSS = 2e-5
TOL_BF16_GRAD = 1e-2


def bf16_round(t):
    return t.bfloat16().float()


def run(pred, tar, mode, dev, gout=None):
    p, t = pred.to(dev), tar.to(dev)
    loss, lse = ops.bbc_loss_forward(p, t, 100.0, mode)
    dp, dt = ops.bbc_loss_backward(p, t, lse, gout, 100.0, mode)
    return float(loss.item()), lse.cpu().numpy(), dp.cpu().numpy(), dt.cpu().numpy()


@pytest.mark.parametrize("dim", [640, 512])
def test_fp32_against_reference_golden(cuda_device, dim):
    z, meta = load_golden(f"bbcloss{dim}")
    pred, tar = syn.loss_pair(meta["seed"], meta["rows"], dim)
    loss, _, dp, dt = run(pred, tar, MODE_FP32, cuda_device)
    assert abs(loss - float(z["loss"])) <= TOL_F32 * abs(float(z["loss"]))
    gmax = float(np.abs(z["dpred"]).max())
    assert float(np.abs(dp - z["dpred"]).max()) <= TOL_F32_GRAD * gmax
    assert float(np.abs(dt - z["dtar"]).max()) <= TOL_F32_GRAD * max(gmax, float(np.abs(z["dtar"]).max()))


@pytest.mark.parametrize("dim", [640, 512])
def test_bf16_against_reference_golden(cuda_device, dim):
    # tensor-core mode vs the fp32 reference itself: the only difference is the bf16 rounding of the operands
    # (|d cos| <= 5e-4 -> |d logit| <= 0.05 at scale 100), so the bound is loose and absolute
    z, meta = load_golden(f"bbcloss{dim}")
    pred, tar = syn.loss_pair(meta["seed"], meta["rows"], dim)
    loss, _, dp, dt = run(pred, tar, MODE_BF16, cuda_device)
    # (the reference's own training runs this loss under fp16 autocast with fp16 LOGITS, ulp 0.0625 at 100)
    assert abs(loss - float(z["loss"])) <= 2e-2
    gmax = float(np.abs(z["dpred"]).max())
    assert float(np.abs(dp - z["dpred"]).max()) <= 0.25 * gmax
    assert float(np.abs(dt - z["dtar"]).max()) <= 0.25 * gmax


@pytest.mark.parametrize("rows", [1, 2, 7, 32, 63, 64, 65, 100, 256, 257, 300, 1024, 1500])
@pytest.mark.parametrize("mode", [MODE_FP32, MODE_BF16])
def test_ragged_batches_against_oracle(cuda_device, rows, mode):
    dim = 640 if rows % 2 else 512
    pred, tar = syn.loss_pair(900 + rows, rows, dim)
    if mode == MODE_BF16:
        ref = orc.bbc_loss(bf16_round(pred), bf16_round(tar))
        tol_l, tol_g = TOL_BF16_LOSS, TOL_BF16_GRAD
    else:
        ref = orc.bbc_loss(pred, tar)
        tol_l, tol_g = TOL_F32, TOL_F32_GRAD
    loss, lse, dp, dt = run(pred, tar, mode, cuda_device)
    # lse and the diagonal logit are both ~100 in fp32 (ulp 7.6e-6): their difference carries that absolute floor
    assert abs(loss - ref[0]) <= tol_l * abs(ref[0]) + 1e-5, (loss, ref[0])
    assert float(np.abs(lse - ref[1]).max()) <= 2e-5 * float(np.abs(ref[1]).max())
    gmax = max(float(np.abs(ref[2]).max()), float(np.abs(ref[3]).max()))
    # fp32 floor: softmax - onehot is formed from logits ~100 (ulp 7.6e-6), then multiplied by scale / B and a
    # feature entry -- the same floor torch's own fp32 cross_entropy backward has
    floor = (100.0 / rows) * 1.6e-5 * float(tar.abs().max())
    assert float(np.abs(dp - ref[2]).max()) <= tol_g * gmax + floor
    assert float(np.abs(dt - ref[3]).max()) <= tol_g * gmax + floor
    if rows == 1:
        assert loss == 0.0 and not dp.any() and not dt.any()      # one class: loss and gradient are exactly zero


@pytest.mark.parametrize("mode", [MODE_FP32, MODE_BF16])
def test_strided_inputs_and_upstream_gradient(cuda_device, mode):
    rows, dim = 200, 512
    pred, tar = syn.loss_pair(77, rows, dim)
    wide_p = torch.zeros(rows, dim + 64, device=cuda_device)
    wide_t = torch.zeros(rows, dim + 128, device=cuda_device)
    wide_p[:, :dim] = pred.to(cuda_device)
    wide_t[:, :dim] = tar.to(cuda_device)
    gout = torch.tensor([1024.0], device=cuda_device)            # what GradScaler.scale(loss).backward() sends
    l1, _, dp1, dt1 = run(pred, tar, mode, cuda_device)
    loss, lse = ops.bbc_loss_forward(wide_p[:, :dim], wide_t[:, :dim], 100.0, mode)
    dp, dt = ops.bbc_loss_backward(wide_p[:, :dim], wide_t[:, :dim], lse, gout, 100.0, mode)
    assert float(loss.item()) == l1                                # deterministic, layout independent
    scale_tol = 0 if mode == MODE_FP32 else 2 ** -7                 # bf16: d(logits) is rounded after the scaling
    assert np.allclose(dp.cpu().numpy(), 1024.0 * dp1, rtol=scale_tol, atol=scale_tol * np.abs(dp1).max() * 1024)
    assert np.allclose(dt.cpu().numpy(), 1024.0 * dt1, rtol=scale_tol, atol=scale_tol * np.abs(dt1).max() * 1024)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_module_is_a_drop_in_for_autograd(cuda_device, precision):
    # the call sequence of run/train/train_fiq.py:124-137 around the criterion: features that require grad,
    # criterion(fusion_feat, target_feat), scaled backward
    rows, dim = 96, 640
    pred, tar = syn.loss_pair(31, rows, dim)
    w = torch.randn(dim, dim, generator=torch.Generator().manual_seed(5)) * 0.05 + torch.eye(dim)

    def step(criterion):
        wp = w.clone().to(cuda_device).requires_grad_(True)
        p = F.normalize(pred.to(cuda_device) @ wp, dim=-1)         # upstream graph stays in torch
        t = tar.clone().to(cuda_device).requires_grad_(True)
        loss = criterion(p, t)
        (loss * 128.0).backward()
        return float(loss), wp.grad.cpu(), t.grad.cpu()

    def torch_criterion(p, t):                                      # losses/loss.py:10-14
        return F.cross_entropy(100 * p @ t.T, torch.arange(p.shape[0], device=p.device))

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        l_ref, gw_ref, gt_ref = step(torch_criterion)
        l_mine, gw, gt = step(BatchBasedClassificationLoss(precision=precision))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    tol_l, tol_g = (2e-5, 1e-3) if precision == "fp32" else (0.15, 0.15)
    assert abs(l_mine - l_ref) <= tol_l * abs(l_ref)
    assert float((gw - gw_ref).abs().max()) <= tol_g * float(gw_ref.abs().max())
    assert float((gt - gt_ref).abs().max()) <= tol_g * float(gt_ref.abs().max())


def test_half_inputs_and_refusals(cuda_device):
    crit = BatchBasedClassificationLoss()
    pred, tar = syn.loss_pair(11, 40, 512)
    p = pred.to(cuda_device).half().requires_grad_(True)           # autocast hands the criterion fp16 features
    t = tar.to(cuda_device).half().requires_grad_(True)
    loss = crit(p, t)
    loss.backward()
    assert loss.dtype == torch.float32 and p.grad.dtype == torch.float16 and t.grad.shape == t.shape
    assert torch.isfinite(p.grad).all() and torch.isfinite(t.grad).all()
    with pytest.raises(ErnError):
        crit(pred, tar)                                             # CPU tensors: no fallback
    with pytest.raises(ErnError):
        crit(pred.to(cuda_device), tar[:10].to(cuda_device))
    with pytest.raises(ErnError):
        ops.bbc_loss_forward(torch.randn(8, 100, device=cuda_device), torch.randn(8, 100, device=cuda_device))
    with pytest.raises(ErnError):
        BatchBasedClassificationLoss(precision="fp8")
