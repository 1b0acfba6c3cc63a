"""Drop-in replacement of the reference's ``VisualSR`` (models/fusion_model.py:97-154) -- SURVEY.md 8(f) row 1.

``VisualSR`` attention-pools the 13 patch embeddings of an image into one vector; it is the step right before
the gallery-side fusion head (``ERN.forward(mode="index")``, models/model.py:64-66) and is used once more inside
``DVR_module`` (models/fusion_model.py:48).  Same constructor, same parameter/buffer names (``embedding_local.0.*``,
``embedding_local.1.*`` incl. BatchNorm running stats, ``embedding_global.*``, ``embedding_common.*``), same Xavier
initialisation (:126-134), eval-mode forward on hand-written sm_100a kernels behind ``ern_visualsr_forward``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib as L
from ._lib import MODE_BF16, MODE_FP32, ErnError, VisualSRWeights
from .ops import launch_counter


class VisualSR(nn.Module):
    def __init__(self, embed_dim=512, dropout_rate=0.5, num_region=13, mode: str = "bf16"):
        super().__init__()
        self.embedding_local = self._layer(embed_dim, num_region, dropout_rate)
        self.embedding_global = self._layer(embed_dim, embed_dim, dropout_rate)
        self.embedding_common = nn.Linear(embed_dim, 1)
        self.softmax = nn.Softmax(dim=1)
        self.dim, self.patches = embed_dim, num_region
        self.mode = mode
        self._packed: Optional[torch.Tensor] = None
        self._folded = None
        self._versions = None
        self.init_weights()

    @staticmethod
    def _layer(embed_dim, num_features, dropout_rate):
        return nn.Sequential(nn.Linear(embed_dim, embed_dim), nn.BatchNorm1d(num_features), nn.Tanh(),
                             nn.Dropout(dropout_rate))

    def init_weights(self):
        """Xavier-uniform Linear weights, zero biases, unit BatchNorm scale (models/fusion_model.py:126-134)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                r = np.sqrt(6.0) / np.sqrt(m.in_features + m.out_features)
                nn.init.uniform_(m.weight, -r, r)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def invalidate_cache(self) -> None:
        """Drop the packed bf16 weight copies.  They are refreshed automatically when a parameter's version counter
        changes (``load_state_dict``, ``copy_``, optimizer steps); call this after writing through ``.data``."""
        self._packed = None
        self._versions = None

    def set_mode(self, mode: str) -> "VisualSR":
        if mode not in ("bf16", "fp32"):
            raise ErnError(f"unknown mode {mode!r}")
        self.mode = mode
        return self

    def _weights(self, device) -> VisualSRWeights:
        lin_l, bn_l = self.embedding_local[0], self.embedding_local[1]
        lin_g, bn_g = self.embedding_global[0], self.embedding_global[1]
        tensors = (lin_l.weight, lin_l.bias, bn_l.weight, bn_l.bias, bn_l.running_mean, bn_l.running_var,
                   lin_g.weight, lin_g.bias, bn_g.weight, bn_g.bias, bn_g.running_mean, bn_g.running_var,
                   self.embedding_common.weight, self.embedding_common.bias)
        for t in tensors:
            if t.device != device or t.dtype != torch.float32:
                raise ErnError("VisualSR parameters must be float32 on the input's CUDA device")
        versions = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._versions != versions:
            with torch.no_grad():
                # eval-mode BatchNorm1d folded to a per-channel affine: y = scale * x + shift (parameter prep only)
                sl = (bn_l.weight / torch.sqrt(bn_l.running_var + bn_l.eps)).contiguous()
                tl = (bn_l.bias - bn_l.running_mean * sl).contiguous()
                sg = (bn_g.weight / torch.sqrt(bn_g.running_var + bn_g.eps)).contiguous()
                tg = (bn_g.bias - bn_g.running_mean * sg).contiguous()
            self._folded = (sl, tl, sg, tg)
            self._packed = None
            self._versions = versions
        sl, tl, sg, tg = self._folded
        w = VisualSRWeights(lin_l.weight.data_ptr(), lin_l.bias.data_ptr(), sl.data_ptr(), tl.data_ptr(),
                            lin_g.weight.data_ptr(), lin_g.bias.data_ptr(), sg.data_ptr(), tg.data_ptr(),
                            self.embedding_common.weight.data_ptr(), self.embedding_common.bias.data_ptr(), None)
        if self.mode == "bf16":
            if self._packed is None or self._packed.device != device:
                packed = torch.empty(L.lib().ern_visualsr_packed_bytes(self.dim), dtype=torch.uint8, device=device)
                L.check(L.lib().ern_visualsr_pack(C.byref(w), self.dim, packed.data_ptr(), L.stream_ptr(device)))
                launch_counter.add(2)
                self._packed = packed
            w.packed_bf16 = self._packed.data_ptr()
        return w

    def forward(self, local_feature: torch.Tensor) -> torch.Tensor:
        """local_feature [B, P, D] -> [B, D] (models/fusion_model.py:141-154), eval mode only."""
        if self.training:
            raise ErnError("VisualSR (B200) implements the eval-mode forward only: call model.eval()")
        if torch.is_grad_enabled() and local_feature.requires_grad:
            raise ErnError("input requires grad: wrap the call in torch.no_grad()")
        L.require_cuda(local_feature, "local_feature")
        if local_feature.dim() != 3 or local_feature.shape[1] != self.patches or local_feature.shape[2] != self.dim:
            raise ErnError(f"expected [B,{self.patches},{self.dim}], got {tuple(local_feature.shape)}")
        x = local_feature.detach().float().contiguous()
        dev, rows = x.device, x.shape[0]
        mode = MODE_BF16 if self.mode == "bf16" else MODE_FP32
        lib = L.lib()
        with torch.cuda.device(dev):
            w = self._weights(dev)
            out = torch.empty((rows, self.dim), dtype=torch.float32, device=dev)
            wsb = lib.ern_visualsr_workspace_bytes(rows, self.patches, self.dim, mode)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
            L.check(lib.ern_visualsr_forward(C.byref(w), self.dim, self.patches, mode, x.data_ptr(), rows,
                                             out.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr(dev)))
            launch_counter.add(4 if rows else 0)
        return out if local_feature.dtype == torch.float32 else out.to(local_feature.dtype)
