"""GPU: the fusion head (ern_combiner_forward) against the reference goldens and the CPU oracle."""
import pytest
import torch

from oracle import ern_oracle as orc
from fashionern_aaai2024_b200 import synthetic as syn
from fashionern_aaai2024_b200.combiner import CombinerSimple
from helpers import load_golden

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-5      # north_star: 1e-5 against the fp32 reference in validation mode
TOL_BF16 = 1e-2      # north_star: 1e-2 relative in bf16 mode


def make(dim, seed, mode, dev):
    m = CombinerSimple(dim, 4 * dim, 8 * dim, mode=mode)
    m.load_state_dict(syn.combiner_state(seed, dim))
    return m.to(dev).eval()


@pytest.mark.parametrize("dim", [640, 512])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_against_reference_golden(cuda_device, dim, mode):
    z, meta = load_golden(f"combiner{dim}")
    m = make(dim, meta["seed"], mode, cuda_device)
    tol = TOL_FP32 if mode == "fp32" else TOL_BF16
    for case in ("raw", "unit", "zero"):
        img = torch.from_numpy(z[f"{case}_image"]).to(cuda_device)
        txt = torch.from_numpy(z[f"{case}_text"]).to(cuda_device)
        with torch.no_grad():
            out = m(img, txt).cpu()
        ref = torch.from_numpy(z[f"{case}_out"])
        err = (out - ref).norm(dim=-1) / ref.norm(dim=-1).clamp_min(1e-12)
        assert float(err.max()) <= tol, (case, float(err.max()))


@pytest.mark.parametrize("rows", [1, 127, 129, 1000])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ragged_batches_against_oracle(cuda_device, rows, mode):
    dim = 512
    sd = syn.combiner_state(77, dim)
    m = make(dim, 77, mode, cuda_device)
    img, txt = syn.features(78, rows, dim), syn.features(79, rows, dim, unit=True)
    ref, gate = orc.combiner_forward(sd, img, txt, return_gate=True)
    with torch.no_grad():
        out, out_b = m(img.to(cuda_device), txt.to(cuda_device), want_bf16=True)
    tol = TOL_FP32 if mode == "fp32" else TOL_BF16
    err = (out.cpu() - ref).norm(dim=-1)
    assert float(err.max()) <= tol
    assert float((out_b.float().cpu() - out.cpu()).abs().max()) <= 2 ** -8      # bf16 rounding of unit vectors
    assert rows < 100 or float(gate.max() - gate.min()) > 0.2                    # the synthetic gate is informative


@pytest.mark.parametrize("dim", [640, 512, 768])
@pytest.mark.parametrize("rows", [1, 15, 16, 17, 31, 32, 33, 48, 64, 65])
def test_small_batches_weight_streaming_path(cuda_device, dim, rows):
    """<= 64 rows take the weight-streaming kernels (ern_combiner_small.cu; the reference's query side runs 32-row
    batches, run/test/test_fiq.py:132); 65 rows is the first size back on the tcgen05 GEMMs.  Same tolerance, both
    outputs, and the result must not depend on which path a row went through."""
    sd = syn.combiner_state(91, dim)
    m = make(dim, 91, "bf16", cuda_device)
    img, txt = syn.features(92, rows, dim), syn.features(93, rows, dim, unit=True)
    ref = orc.combiner_forward(sd, img, txt)
    with torch.no_grad():
        out, out_b = m(img.to(cuda_device), txt.to(cuda_device), want_bf16=True)
        big = m(torch.cat([img, syn.features(94, 200, dim)]).to(cuda_device),
                torch.cat([txt, syn.features(95, 200, dim, unit=True)]).to(cuda_device))[:rows]
    assert float((out.cpu() - ref).norm(dim=-1).max()) <= TOL_BF16
    assert float((out_b.float().cpu() - out.cpu()).abs().max()) <= 2 ** -8
    # small path vs tcgen05 path on the same rows: both are bf16 operands with fp32 accumulation
    assert float((out - big).norm(dim=-1).max()) <= 2e-3
    assert torch.isfinite(out).all()


def test_small_batch_head_on_concurrent_streams(cuda_device):
    """The fused small-batch kernel synchronises its CTAs through counters that live behind the packed weights (15
    rotating, self-resetting sets): forwards of ONE module issued on several streams at once must neither hang nor mix
    results."""
    dim = 640
    m = make(dim, 33, "bf16", cuda_device)
    xs = [(syn.features(100 + i, 32, dim).to(cuda_device), syn.features(200 + i, 32, dim, unit=True).to(cuda_device))
          for i in range(4)]
    with torch.no_grad():
        want = [m(a, b).clone() for a, b in xs]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(cuda_device) for _ in range(4)]
        outs = [[] for _ in range(4)]
        for rep in range(12):
            for i, st in enumerate(streams):
                with torch.cuda.stream(st):
                    outs[i].append(m(*xs[i]))
        torch.cuda.synchronize()
    for i in range(4):
        for o in outs[i]:
            assert torch.equal(o, want[i])


def test_dvr_call_site_inputs(cuda_device):
    # the three query-side call sites (models/fusion_model.py:52-54) with the inputs the reference recorded
    z, meta = load_golden("fiq640")
    for i, name in enumerate(("DVR.combiner_global", "DVR.combiner_local", "DVR.combiner")):
        m = make(640, meta["seed"] + 10 + i, "fp32", cuda_device)
        with torch.no_grad():
            out = m(torch.from_numpy(z[f"io_{name}_image"]).to(cuda_device),
                    torch.from_numpy(z[f"io_{name}_text"]).to(cuda_device)).cpu()
        assert float((out - torch.from_numpy(z[f"io_{name}_out"])).abs().max()) <= TOL_FP32


def test_weight_update_invalidates_packed_cache(cuda_device):
    m = make(512, 5, "bf16", cuda_device)
    x, y = syn.features(1, 8, 512).to(cuda_device), syn.features(2, 8, 512).to(cuda_device)
    a = m(x, y)
    m.load_state_dict(syn.combiner_state(6, 512))
    b = m(x, y)
    ref = orc.combiner_forward(syn.combiner_state(6, 512), x.cpu(), y.cpu())
    assert float((b.cpu() - ref).norm(dim=-1).max()) <= TOL_BF16 and not torch.equal(a, b)
