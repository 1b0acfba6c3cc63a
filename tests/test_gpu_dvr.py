"""GPU: the query-side DVR encoder against the reference golden and the CPU oracle."""
import pytest
import torch

from oracle import ern_oracle as orc
from fashionern_aaai2024_b200 import synthetic as syn
from fashionern_aaai2024_b200.dvr import DVR_module
from helpers import load_golden

pytestmark = pytest.mark.gpu

# fp32 validation mode: 2 BERT layers + cross attention + 4 heads of fp32 sums in a different order than MKL
TOL = {"fp32": 2e-5, "bf16": 1e-2}


def make(dim, seed, mode, dev):
    m = DVR_module(dim, mode=mode)
    m.load_state_dict(syn.dvr_full_state(seed, dim))
    return m.to(dev).eval()


def inputs(seed, rows, dim, dev):
    return [t.to(dev) for t in (syn.patch_features(seed + 10, rows, dim), syn.token_features(seed + 11, rows, dim),
                                syn.features(seed + 12, rows, dim), syn.features(seed + 13, rows, dim))]


@pytest.mark.parametrize("dim", [640, 512])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_against_reference_golden(cuda_device, dim, mode):
    z, meta = load_golden(f"dvr{dim}")
    m = make(dim, meta["seed"], mode, cuda_device)
    with torch.no_grad():
        out = m(*inputs(meta["seed"], meta["rows"], dim, cuda_device)).cpu()
    err = (out - torch.from_numpy(z["out"])).norm(dim=-1)
    assert float(err.max()) <= TOL[mode], float(err.max())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_encoder_outputs_against_oracle(cuda_device, mode):
    dim, rows, seed = 640, 37, 700
    sd = syn.dvr_full_state(seed, dim)
    m = make(dim, seed, mode, cuda_device)
    m.max_batch = 16                                       # exercise the chunk loop (16 + 16 + 5)
    patches, tokens, ref_g, txt_g = inputs(seed, rows, dim, cuda_device)
    with torch.no_grad():
        out = m(patches, tokens, ref_g, txt_g).cpu()
        cross, seq_mean = m.encode(patches, tokens)
    ref = orc.dvr_forward(sd, patches.cpu(), tokens.cpu(), ref_g.cpu(), txt_g.cpu())
    assert float((out - ref).norm(dim=-1).max()) <= TOL[mode]
    assert cross.shape == (rows, 13, dim) and seq_mean.shape == (rows, dim)
    assert bool(torch.isfinite(cross).all()) and float(seq_mean.norm(dim=-1).max()) <= 1.0 + 1e-4


def test_state_dict_matches_reference_names(cuda_device):
    m = DVR_module(512)
    keys = set(m.state_dict().keys())
    assert keys == set(syn.dvr_full_state(1, 512).keys())
    sd = syn.dvr_full_state(1, 512)
    sd["transformer_layer.bert_encoder.bert_model.embeddings.position_ids"] = torch.arange(512)[None]   # old HF buffer
    del sd["transformer_layer.cls_token"]                                                                # CUDA-built ckpt
    m.load_state_dict(sd)
