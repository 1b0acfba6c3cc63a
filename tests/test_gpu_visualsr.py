"""GPU: VisualSR against the reference golden and the CPU oracle."""
import pytest
import torch

from oracle import ern_oracle as orc
from fashionern_aaai2024_b200 import synthetic as syn
from fashionern_aaai2024_b200.visual_sr import VisualSR
from helpers import load_golden

pytestmark = pytest.mark.gpu

# "bf16" = the 16-bit tensor-core mode; VisualSR feeds its GEMMs fp16 operands there (the 13-way softmax amplifies operand
# rounding: with bf16 operands the worst row of a random case reached 1.06e-2), observed worst 1.1e-3
TOL = {"fp32": 1e-5, "bf16": 2.5e-3}


def make(dim, seed, mode, dev):
    m = VisualSR(dim, mode=mode)
    m.load_state_dict(syn.visualsr_state(seed, dim))
    return m.to(dev).eval()


@pytest.mark.parametrize("dim", [640, 512])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_against_reference_golden(cuda_device, dim, mode):
    z, meta = load_golden(f"visualsr{dim}")
    m = make(dim, meta["seed"], mode, cuda_device)
    x = syn.patch_features(meta["seed"] + 1, meta["rows"], dim).to(cuda_device)
    with torch.no_grad():
        out = m(x).cpu()
    err = (out - torch.from_numpy(z["out"])).norm(dim=-1)
    assert float(err.max()) <= TOL[mode], float(err.max())


@pytest.mark.parametrize("rows", [1, 9, 10, 333, 4000])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ragged_batches_against_oracle(cuda_device, rows, mode):
    dim = 640
    sd = syn.visualsr_state(55, dim)
    m = make(dim, 55, mode, cuda_device)
    x = syn.patch_features(56, rows, dim)
    ref = orc.visual_sr_forward(sd, x)
    with torch.no_grad():
        out = m(x.to(cuda_device)).cpu()
    assert float((out - ref).norm(dim=-1).max()) <= TOL[mode]


def test_state_dict_keys_and_refusals(cuda_device):
    m = VisualSR(512)
    assert {"embedding_local.0.weight", "embedding_local.1.running_var", "embedding_global.1.weight",
            "embedding_common.bias", "embedding_local.1.num_batches_tracked"} <= set(m.state_dict().keys())
    m = m.to(cuda_device)
    with pytest.raises(Exception):
        m(torch.randn(2, 13, 512, device=cuda_device))          # training mode refused
    with pytest.raises(Exception):
        m.eval()(torch.randn(2, 13, 512))                       # CPU tensor refused
