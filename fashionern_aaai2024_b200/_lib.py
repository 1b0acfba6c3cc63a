"""ctypes binding of ``libern_b200.so`` (C ABI declared in ``include/ern_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile`` and must be
present: there is NO CPU or PyTorch fallback for any operation of this package.  Loading fails loudly
when the library is missing, and every compute call fails loudly (``ErnError``) on a device that is not
sm_100-class.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ERN_B200_LIB: load another build of the same C ABI (A/B timing of two builds in one GPU session); default in-tree
LIB_PATH = os.environ.get("ERN_B200_LIB") or os.path.join(_HERE, "libern_b200.so")

MODE_BF16, MODE_FP32 = 0, 1
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
NORM_OUT_F16 = 2          # ern_l2norm_rows: OR into `normalize` for an fp16 (instead of bf16) 16-bit output
RANK_SIMILARITY, RANK_REFERENCE = 0, 1
MAX_K, SEG_CAP, SORT_CAP, QUERY_BATCH, DENSE_ROWS = 128, 512, 2048, 4096, 256

# every symbol include/ern_b200.h declares; tests check the .so exports exactly these
SYMBOLS = (
    "ern_version", "ern_last_error", "ern_device_check", "ern_l2norm_rows",
    "ern_combiner_packed_bytes", "ern_combiner_pack", "ern_combiner_workspace_bytes", "ern_combiner_forward",
    "ern_sim_topk_workspace_bytes", "ern_sim_topk", "ern_sim_topk_exchange", "ern_topk_merge", "ern_recall_at_k",
    "ern_cirr_subset_recall", "ern_gather_scores", "ern_cirr_subset_from_scores",
    "ern_visualsr_packed_bytes", "ern_visualsr_pack", "ern_visualsr_workspace_bytes", "ern_visualsr_forward",
    "ern_dvr_packed_bytes", "ern_dvr_pack", "ern_dvr_workspace_bytes", "ern_dvr_encode",
    "ern_bbc_loss_workspace_bytes", "ern_bbc_loss_forward", "ern_bbc_loss_backward",
)


class ErnError(RuntimeError):
    pass


class CombinerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_text", "b_text", "w_image", "b_image", "w_hid", "b_hid",
                                          "w_gate", "b_gate", "packed_bf16")]


class VisualSRWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_local", "b_local", "bn_local_scale", "bn_local_shift", "w_global",
                                          "b_global", "bn_global_scale", "bn_global_shift", "w_common", "b_common",
                                          "packed_bf16")]


class BertLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "ln1_w", "ln1_b", "wi", "bi",
                                          "wo2", "bo2", "ln2_w", "ln2_b")]


class DvrWeights(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("cls_token", "pos_emb", "type_emb", "emb_ln_w", "emb_ln_b")]
                + [("n_layers", C.c_int), ("intermediate", C.c_int), ("layers", BertLayerWeights * 4)]
                + [(n, C.c_void_p) for n in ("mha_in_w", "mha_in_b", "mha_out_w", "mha_out_b", "packed_bf16")])


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ErnError(
            f"{LIB_PATH} not found. Build it first (python -c 'import __graft_entry__ as g; g.build()' or "
            f"make -C {os.path.join(_HERE, 'csrc')}). This package has no CPU fallback.")
    l = C.CDLL(LIB_PATH)
    i64, i32, vp, sz = C.c_int64, C.c_int, C.c_void_p, C.c_size_t
    l.ern_version.restype = i32
    l.ern_last_error.restype = C.c_char_p
    l.ern_device_check.argtypes = [i32]
    l.ern_l2norm_rows.argtypes = [vp, i64, i32, i64, i32, vp, i64, vp, i64, vp]
    l.ern_combiner_packed_bytes.argtypes = [i32]
    l.ern_combiner_packed_bytes.restype = sz
    l.ern_combiner_pack.argtypes = [C.POINTER(CombinerWeights), i32, vp, vp]
    l.ern_combiner_workspace_bytes.argtypes = [i64, i32, i32]
    l.ern_combiner_workspace_bytes.restype = sz
    l.ern_combiner_forward.argtypes = [C.POINTER(CombinerWeights), i32, i32, vp, vp, i64, vp, vp, i64, vp, vp, sz, vp]
    l.ern_sim_topk_workspace_bytes.argtypes = [i64, i32, i32]
    l.ern_sim_topk_workspace_bytes.restype = sz
    l.ern_sim_topk.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i64, vp, i32, i32, i32, i32,
                               vp, vp, vp, vp, vp, sz, vp]
    l.ern_sim_topk_exchange.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i64, vp, i32, i32, i32, i32,
                                        vp, i32, i32, vp, vp, sz, vp]
    l.ern_topk_merge.argtypes = [vp, i64, i32, i32, i64, i64, i32, vp, vp, vp, vp]
    l.ern_recall_at_k.argtypes = [vp, i64, i32, vp, i64, vp, C.POINTER(C.c_int32), i32, vp, vp, vp]
    l.ern_cirr_subset_recall.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, vp, i32, vp, vp, i32,
                                         C.POINTER(C.c_int32), i32, vp, vp, vp]
    l.ern_gather_scores.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i64, vp, i32, vp, vp]
    l.ern_cirr_subset_from_scores.argtypes = [vp, i64, vp, i32, vp, vp, i32, C.POINTER(C.c_int32), i32, vp, vp, vp]
    l.ern_bbc_loss_workspace_bytes.argtypes = [i64, i32, i32]
    l.ern_bbc_loss_workspace_bytes.restype = sz
    l.ern_bbc_loss_forward.argtypes = [vp, i64, vp, i64, i64, i32, C.c_float, i32, vp, vp, vp, sz, vp]
    l.ern_bbc_loss_backward.argtypes = [vp, i64, vp, i64, i64, i32, C.c_float, i32, vp, vp, vp, i64, vp, i64, vp, sz, vp]
    l.ern_visualsr_packed_bytes.argtypes = [i32]
    l.ern_visualsr_packed_bytes.restype = sz
    l.ern_visualsr_pack.argtypes = [C.POINTER(VisualSRWeights), i32, vp, vp]
    l.ern_visualsr_workspace_bytes.argtypes = [i64, i32, i32, i32]
    l.ern_visualsr_workspace_bytes.restype = sz
    l.ern_visualsr_forward.argtypes = [C.POINTER(VisualSRWeights), i32, i32, i32, vp, i64, vp, vp, sz, vp]
    l.ern_dvr_packed_bytes.argtypes = [i32, i32, i32]
    l.ern_dvr_packed_bytes.restype = sz
    l.ern_dvr_pack.argtypes = [C.POINTER(DvrWeights), i32, vp, vp]
    l.ern_dvr_workspace_bytes.argtypes = [i64, i32, i32, i32, i32, i32]
    l.ern_dvr_workspace_bytes.restype = sz
    l.ern_dvr_encode.argtypes = [C.POINTER(DvrWeights), i32, i32, i32, i32, i32, vp, vp, i64, vp, vp, vp, sz, vp]
    for name in SYMBOLS:
        fn = getattr(l, name)
        if fn.restype is C.c_int and name not in ("ern_version",):
            fn.restype = i32
    _lib = l
    return l


def check(rc: int) -> None:
    if rc != 0:
        raise ErnError(f"libern_b200 error {rc}: {lib().ern_last_error().decode()}")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise ErnError(f"{what} must be a CUDA tensor (got device {t.device}); this package has no CPU fallback")


def require_device(device: torch.device) -> None:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(lib().ern_device_check(idx))
