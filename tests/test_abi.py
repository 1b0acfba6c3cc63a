"""CPU: the C-ABI library loads, exports exactly what include/ern_b200.h declares, and the product
fails loudly (no CPU fallback) when there is no sm_100 device or a CPU tensor is passed."""
import ctypes
import os
import re

import pytest
import torch

from fashionern_aaai2024_b200 import _lib, ops
from fashionern_aaai2024_b200.combiner import CombinerSimple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "ern_b200.h")) as f:
        src = f.read()
    return re.findall(r"ERN_API\s+[\w\s\*]+?\b(ern_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    declared = header_symbols()
    assert sorted(declared) == sorted(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_constants():
    lib = _lib.lib()
    assert lib.ern_version() >= 100
    with open(os.path.join(ROOT, "include", "ern_b200.h")) as f:
        src = f.read()
    assert int(re.search(r"#define ERN_MAX_K (\d+)", src).group(1)) == _lib.MAX_K
    assert int(re.search(r"#define ERN_SEG_CAP (\d+)", src).group(1)) == _lib.SEG_CAP
    assert int(re.search(r"#define ERN_SORT_CAP (\d+)", src).group(1)) == _lib.SORT_CAP
    assert int(re.search(r"#define ERN_DENSE_ROWS (\d+)", src).group(1)) == _lib.DENSE_ROWS
    assert int(re.search(r"#define ERN_QUERY_BATCH (\d+)", src).group(1)) == _lib.QUERY_BATCH
    for name in ("DTYPE_F32", "DTYPE_BF16", "DTYPE_F16", "NORM_OUT_F16", "MODE_BF16", "MODE_FP32", "RANK_SIMILARITY",
                 "RANK_REFERENCE"):
        assert int(re.search(rf"#define ERN_{name} (\d+)", src).group(1)) == getattr(_lib, name), name
    # candidate storage of one query batch: a 256-slot prefix + one 256-slot segment per CTA pair (74 on a B200)
    one_batch = lib.ern_sim_topk_workspace_bytes(4096, 640, 0)
    assert one_batch >= 4096 * (_lib.DENSE_ROWS + 74 * _lib.SEG_CAP) * 8
    # larger calls are processed one batch at a time in the same storage
    assert lib.ern_sim_topk_workspace_bytes(33480, 640, 0) == one_batch
    assert lib.ern_combiner_packed_bytes(640) >= (8 * 640 * 640 + 64 * 640 * 640) * 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    rc = _lib.lib().ern_device_check(0)
    assert rc != 0 and len(_lib.lib().ern_last_error()) > 0
    with pytest.raises(_lib.ErnError):
        _lib.check(rc)


def test_cpu_tensors_are_refused():
    x = torch.randn(4, 64)
    with pytest.raises(_lib.ErnError):
        ops.l2norm_rows(x)
    with pytest.raises(_lib.ErnError):
        ops.sim_topk(x, x, 2, mode=_lib.MODE_FP32)
    m = CombinerSimple(64, 256, 512).eval()
    with pytest.raises(_lib.ErnError):
        m(x, x)


def test_feature_dtype_rules_of_the_scoring_call():
    # host-side argument checks of ops.sim_topk and friends (they run before any device work): both matrices share one
    # type; fp32 validation mode takes float32, the tensor-core mode bfloat16 or float16 (ERN_DTYPE_F16)
    from fashionern_aaai2024_b200 import metrics
    f32, bf, h = torch.zeros(2, 64), torch.zeros(2, 64, dtype=torch.bfloat16), torch.zeros(2, 64, dtype=torch.float16)
    assert ops._feature_dtype(f32, f32, _lib.MODE_FP32) == _lib.DTYPE_F32
    assert ops._feature_dtype(bf, bf, _lib.MODE_BF16) == _lib.DTYPE_BF16
    assert ops._feature_dtype(h, h, _lib.MODE_BF16) == _lib.DTYPE_F16
    assert ops._feature_dtype(h, h) == _lib.DTYPE_F16                   # (calls without a mode: subset recall, gathers)
    for q, g, mode in ((h, bf, _lib.MODE_BF16), (f32, f32, _lib.MODE_BF16), (h, h, _lib.MODE_FP32), (bf, bf, _lib.MODE_FP32),
                       (f32.double(), f32.double(), _lib.MODE_FP32), (f32, bf, None)):
        with pytest.raises(_lib.ErnError):
            ops._feature_dtype(q, g, mode)
    for p in ("bf16", "fp16", "fp32"):
        metrics.set_precision(p)
    metrics.set_precision("bf16")
    with pytest.raises(_lib.ErnError):
        metrics.set_precision("fp8")
    with pytest.raises(_lib.ErnError):
        metrics._operands(f32, f32, "int8")


def test_training_mode_is_refused():
    m = CombinerSimple(64, 256, 512)
    m.train()
    with pytest.raises(_lib.ErnError):
        m(torch.randn(2, 64), torch.randn(2, 64))


def test_state_dict_keys_match_reference():
    # models/fusion_model.py:73-84 -> the key names ERN.load_state_dict expects (run/test/test_fiq.py:149)
    keys = set(CombinerSimple(64, 256, 512).state_dict().keys())
    assert keys == {
        "dynamic_scalar.0.weight", "dynamic_scalar.0.bias", "dynamic_scalar.3.weight", "dynamic_scalar.3.bias",
        "text_projection_layer.0.weight", "text_projection_layer.0.bias",
        "image_projection_layer.0.weight", "image_projection_layer.0.bias"}
    sd = CombinerSimple(64, 256, 512).state_dict()
    assert sd["dynamic_scalar.0.weight"].shape == (512, 512) and sd["dynamic_scalar.3.weight"].shape == (1, 512)
    assert sd["text_projection_layer.0.weight"].shape == (256, 64)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference checkout (authoring container)")
def test_accelerate_ern_swaps_reference_modules_and_keeps_weights():
    from oracle import ref_harness as ref
    from fashionern_aaai2024_b200 import VisualSR, accelerate_ern, synthetic as syn
    states = {n: syn.combiner_state(20 + i, 64) for i, n in enumerate(
        ("DVR.combiner_global", "DVR.combiner_local", "DVR.combiner", "Combiner_module"))}
    model = ref.build_ern(64, states)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    model = accelerate_ern(model, mode="fp32")
    assert isinstance(model.Combiner_module, CombinerSimple) and isinstance(model.DVR.combiner, CombinerSimple)
    from fashionern_aaai2024_b200 import DVR_module
    assert isinstance(model.DVR, DVR_module)
    assert isinstance(model.SR_module, VisualSR) and isinstance(model.DVR.SR_module, VisualSR)
    after = model.state_dict()
    assert set(after.keys()) == set(before.keys())             # checkpoints stay loadable both ways
    assert all(torch.equal(after[k], before[k]) for k in before)
    assert not model.Combiner_module.training                   # eval flag preserved


def test_ern_checkpoint_carries_and_restores_clip_weights():
    """The reference registers the CLIP model as a submodule of ImageCLIP / TextCLIP (models/clip_model.py:8,21), so
    its fusion checkpoints carry the backbone under image_clip.clip_model.* / text_clip.clip_model.* and
    load_state_dict restores it (run/test/test_fiq.py:148-149).  The B200 ERN keeps the backbone outside its own
    parameters but must save and restore those keys the same way, and must say so when it cannot."""
    import warnings
    import fashionern_aaai2024_b200 as ern

    class Clip(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 4)

    src_clip = Clip()
    sd = ern.ERN(src_clip, 64, None).state_dict()
    clip_keys = sorted(k for k in sd if "clip" in k)
    assert clip_keys == ["image_clip.clip_model.lin.bias", "image_clip.clip_model.lin.weight",
                         "text_clip.clip_model.lin.bias", "text_clip.clip_model.lin.weight"]
    dst_clip = Clip()
    assert not torch.equal(dst_clip.lin.weight, src_clip.lin.weight)
    ern.ERN(dst_clip, 64, None).load_state_dict(sd)
    assert torch.equal(dst_clip.lin.weight, src_clip.lin.weight) and torch.equal(dst_clip.lin.bias, src_clip.lin.bias)
    # a stand-in that is not an nn.Module cannot take them: loud warning, the rest still loads
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ern.ERN(object(), 64, None).load_state_dict(sd)
    assert any("CLIP backbone weights" in str(x.message) for x in w)


def test_header_is_plain_c_and_links(tmp_path):
    """include/ern_b200.h must be consumable by a C compiler (the boundary is a C ABI, not C++) and every declared
    function must resolve against the built library."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    calls = "\n".join(f"  p[{i}] = (fn_t)&{name};" for i, name in enumerate(_lib.SYMBOLS))
    src = tmp_path / "abi.c"
    src.write_text('#include "ern_b200.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t p[%d];\n%s\n'
                   '  printf("%%d %%d\\n", ern_version(), (int)sizeof(ern_dvr_weights));\n  return p[0] == 0;\n}\n'
                   % (len(_lib.SYMBOLS), calls))
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-l:libern_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) >= 100
    # the ctypes mirror of the largest struct must have the C layout
    import ctypes
    assert int(out[1]) == ctypes.sizeof(_lib.DvrWeights)


@pytest.mark.gpu
def test_plain_c_caller_end_to_end(tmp_path, cuda_device):
    """examples/ern_topk_demo.c: a C99 program drives ern_sim_topk with cudaMalloc'ed buffers -- no Python, no C++."""
    import shutil
    import subprocess
    cuda = "/usr/local/cuda"
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no gcc / CUDA headers")
    exe = tmp_path / "demo"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(ROOT, "examples", "ern_topk_demo.c"), "-o", str(exe), "-L", libdir, "-l:libern_b200.so",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{libdir}",
                    f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
