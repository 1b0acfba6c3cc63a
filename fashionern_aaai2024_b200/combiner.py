"""Drop-in replacement of the reference's fusion head ``CombinerSimple`` (models/fusion_model.py:58-94).

Same constructor, same parameter / ``state_dict`` key names (so ``ERN.load_state_dict`` of a reference
checkpoint works unchanged, run/test/test_fiq.py:149), same ``forward(image_features, text_features)``
argument order and semantics -- but the forward is one chain of hand-written sm_100a kernels behind the
C ABI (``ern_combiner_forward``): the projections, the hidden layer, the gate, the blend and the final
L2-normalise never round-trip through torch ops.

Only inference is on the accelerated path (the reference calls the head under ``model.eval()`` +
``torch.no_grad()``, run/test/test_fiq.py:100,168).  Training-mode forward (active Dropout, autograd)
is refused loudly rather than silently diverging; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib as L
from ._lib import MODE_BF16, MODE_FP32, CombinerWeights, ErnError
from .ops import launch_counter


class CombinerSimple(nn.Module):
    """Gated dynamic-scalar combiner: ``normalize(s * text + (1 - s) * image)`` with
    ``s = sigmoid(W2 relu(W1 [relu(Wt text + bt) | relu(Wi image + bi)] + b1) + b2)``."""

    def __init__(self, clip_feature_dim=512, projection_dim=512 * 4, hidden_dim=512 * 8, mode: str = "bf16"):
        super().__init__()
        if projection_dim != 4 * clip_feature_dim or hidden_dim != 8 * clip_feature_dim:
            raise ErnError("the accelerated head supports the reference's (D, 4D, 8D) configuration only "
                           "(models/model.py:20, models/fusion_model.py:22-24)")
        # identical layer structure => identical parameter names as models/fusion_model.py:73-84
        self.dynamic_scalar = nn.Sequential(
            nn.Linear(projection_dim * 2, hidden_dim), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(hidden_dim, 1), nn.Sigmoid())
        self.text_projection_layer = nn.Sequential(nn.Linear(clip_feature_dim, projection_dim), nn.ReLU(), nn.Dropout(0.5))
        self.image_projection_layer = nn.Sequential(nn.Linear(clip_feature_dim, projection_dim), nn.ReLU(), nn.Dropout(0.5))
        self.dim = clip_feature_dim
        self.set_mode(mode)
        self._packed: Optional[torch.Tensor] = None
        self._packed_versions = None

    # ---------------------------------------------------------------------------------------------
    def invalidate_cache(self) -> None:
        """Drop the packed bf16 weight copies.  They are refreshed automatically when a parameter's version counter
        changes (``load_state_dict``, ``copy_``, optimizer steps); call this after writing through ``.data``."""
        self._packed = None
        self._packed_versions = None

    def set_mode(self, mode: str) -> "CombinerSimple":
        """'bf16' = tcgen05 tensor-core path (default); 'fp32' = FFMA validation mode (1e-5 vs the reference)."""
        if mode not in ("bf16", "fp32"):
            raise ErnError(f"unknown mode {mode!r}")
        self.mode = mode
        return self

    def _params(self):
        return (self.text_projection_layer[0].weight, self.text_projection_layer[0].bias,
                self.image_projection_layer[0].weight, self.image_projection_layer[0].bias,
                self.dynamic_scalar[0].weight, self.dynamic_scalar[0].bias,
                self.dynamic_scalar[3].weight, self.dynamic_scalar[3].bias)

    def _weights(self, device) -> CombinerWeights:
        ps = self._params()
        for p in ps:
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise ErnError("combiner parameters must be contiguous float32 on the input's CUDA device "
                               "(call model.float().to(device) as the reference does, run/test/test_fiq.py:148,169)")
        w = CombinerWeights(*[p.data_ptr() for p in ps], None)
        if self.mode == "bf16":
            versions = tuple((p.data_ptr(), p._version) for p in ps)
            if self._packed is None or self._packed_versions != versions or self._packed.device != device:
                # bf16 K-major copies of the three weight matrices, refreshed whenever a parameter changes
                nbytes = L.lib().ern_combiner_packed_bytes(self.dim)
                packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
                L.check(L.lib().ern_combiner_pack(C.byref(w), self.dim, packed.data_ptr(), L.stream_ptr(device)))
                launch_counter.add(3)
                self._packed, self._packed_versions = packed, versions
            w.packed_bf16 = self._packed.data_ptr()
        return w

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor, want_bf16: bool = False):
        """(image_features [B,D], text_features [B,D]) -> fused unit-norm features [B,D] float32
        (models/fusion_model.py:86-94).  ``want_bf16=True`` additionally returns the bf16 rounding that
        the scoring kernel consumes, produced by the same epilogue."""
        if self.training:
            raise ErnError("CombinerSimple (B200) implements the eval-mode forward only: call model.eval(). "
                           "Training-mode Dropout/autograd is outside the accelerated path")
        if torch.is_grad_enabled() and (image_features.requires_grad or text_features.requires_grad):
            raise ErnError("inputs require grad: the accelerated head has no backward; wrap the call in torch.no_grad()")
        L.require_cuda(image_features, "image_features")
        L.require_cuda(text_features, "text_features")
        if image_features.shape != text_features.shape or image_features.dim() != 2 or image_features.shape[1] != self.dim:
            raise ErnError(f"expected two [B,{self.dim}] tensors, got {tuple(image_features.shape)} and {tuple(text_features.shape)}")
        out_dtype = image_features.dtype
        img = image_features.detach().float().contiguous()
        txt = text_features.detach().float().contiguous()
        dev = img.device
        rows = img.shape[0]
        mode = MODE_BF16 if self.mode == "bf16" else MODE_FP32
        lib = L.lib()
        with torch.cuda.device(dev):
            w = self._weights(dev)
            out = torch.empty((rows, self.dim), dtype=torch.float32, device=dev)
            out_b = torch.empty((rows, self.dim), dtype=torch.bfloat16, device=dev) if want_bf16 else None
            wsb = lib.ern_combiner_workspace_bytes(rows, self.dim, mode)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
            L.check(lib.ern_combiner_forward(C.byref(w), self.dim, mode, img.data_ptr(), txt.data_ptr(), rows,
                                             out.data_ptr(), L.ptr(out_b), self.dim, None, ws.data_ptr(), wsb,
                                             L.stream_ptr(dev)))
            # bf16: 2 casts, 2 projection GEMMs, gate GEMM, finaliser (<= 64 rows: ONE cooperative weight-streaming launch);
            # fp32: 2 projections, hidden+gate, finaliser
            small = mode == MODE_BF16 and rows <= 64 and self.dim % 64 == 0
            launch_counter.add(0 if not rows else (1 if small else 6 if mode == MODE_BF16 else 4))
        if out_dtype != torch.float32:
            out = out.to(out_dtype)
        return (out, out_b) if want_bf16 else out


def accelerate_ern(model: nn.Module, mode: str = "bf16") -> nn.Module:
    """Swap every reference ``CombinerSimple`` inside an ``ERN`` (models/model.py:16-20: ``DVR.combiner_global``,
    ``DVR.combiner_local``, ``DVR.combiner``, ``Combiner_module``) and every reference ``VisualSR`` (``SR_module``,
    ``DVR.SR_module``) and the whole query-side ``DVR_module`` (``DVR``) for the B200 modules, keeping their weights
    and buffers."""
    from .dvr import DVR_module
    from .visual_sr import VisualSR
    for name, child in list(model.named_children()):
        cls = type(child).__name__
        if cls == "DVR_module" and not isinstance(child, DVR_module) and hasattr(child, "MR_component"):
            dim = child.MR_component.embed_dim
            new = DVR_module(dim, device=None, layers=len(child.transformer_layer.bert_encoder.bert_model.encoder.layer),
                             mode=mode)
        elif cls == "CombinerSimple" and not isinstance(child, CombinerSimple) and hasattr(child, "dynamic_scalar"):
            dim = child.text_projection_layer[0].in_features
            new = CombinerSimple(dim, dim * 4, dim * 8, mode=mode)
        elif cls == "VisualSR" and not isinstance(child, VisualSR) and hasattr(child, "embedding_common"):
            dim = child.embedding_common.in_features
            new = VisualSR(dim, num_region=child.embedding_local[1].num_features, mode=mode)
        else:
            accelerate_ern(child, mode)
            continue
        new.load_state_dict(child.state_dict())
        new = new.to(next(child.parameters()).device).float()
        new.train(child.training)
        setattr(model, name, new)
    return model
