"""Drop-in replacement of the reference's ``DVR_module`` (models/fusion_model.py:8-55) -- SURVEY.md 8(f) row 2.

``DVR_module`` is the query side of ERN (``model(mode="test")``, models/model.py:68-69): a 2-layer BERT over
[CLS] + 13 patch + 77 token embeddings (``PlusModel``, :157-216), a cross attention from the normalised token states
to the normalised patch states (:38-47), ``VisualSR`` on its first 13 outputs (:48), the mean of the normalised token
states (:49) and three fusion heads (:52-54).  This module keeps the constructor, the attribute tree and therefore
every ``state_dict`` key of the reference (``transformer_layer.bert_encoder.bert_model.encoder.layer.0.attention.self.
query.weight`` ... ``MR_component.in_proj_weight`` ... ``combiner.dynamic_scalar.0.weight``) but holds the BERT
parameters in plain ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Embedding`` containers (no dependency on ``transformers``)
and runs the eval-mode forward on the sm_100a kernels behind ``ern_dvr_encode`` + the B200 ``VisualSR`` and
``CombinerSimple`` modules.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib as L
from ._lib import MODE_BF16, MODE_FP32, BertLayerWeights, DvrWeights, ErnError
from .combiner import CombinerSimple
from .ops import launch_counter
from .visual_sr import VisualSR

_INTERMEDIATE = 3072      # BertConfig default; models/fusion_model.py:162-170 does not override it
_MAX_POS = 512            # max_position_embeddings (:165)


def _ns(**kw) -> nn.Module:
    m = nn.Module()
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def _bert_layer(d: int) -> nn.Module:
    attn_self = _ns(query=nn.Linear(d, d), key=nn.Linear(d, d), value=nn.Linear(d, d))
    attn_out = _ns(dense=nn.Linear(d, d), LayerNorm=nn.LayerNorm(d, eps=1e-12))
    attention = nn.Module()
    attention.add_module("self", attn_self)
    attention.add_module("output", attn_out)
    return _ns(attention=attention, intermediate=_ns(dense=nn.Linear(d, _INTERMEDIATE)),
               output=_ns(dense=nn.Linear(_INTERMEDIATE, d), LayerNorm=nn.LayerNorm(d, eps=1e-12)))


class DVR_module(nn.Module):  # noqa: N801  (the reference's class name)
    def __init__(self, feature_dim=640, device=None, layers: int = 2, mode: str = "bf16"):
        super().__init__()
        d = feature_dim
        self.device, self.dim, self.heads, self.n_layers = device, d, 8, layers
        embeddings = _ns(position_embeddings=nn.Embedding(_MAX_POS, d), token_type_embeddings=nn.Embedding(2, d),
                         LayerNorm=nn.LayerNorm(d, eps=1e-12))
        bert_model = _ns(embeddings=embeddings, encoder=_ns(layer=nn.ModuleList([_bert_layer(d) for _ in range(layers)])),
                         pooler=_ns(dense=nn.Linear(d, d)))
        self.transformer_layer = _ns(bert_encoder=_ns(bert_model=bert_model))
        self.transformer_layer.cls_token = nn.Parameter(torch.zeros(1, 1, d))
        self.SR_module = VisualSR(embed_dim=d, mode=mode)
        self.MR_component = nn.MultiheadAttention(embed_dim=d, num_heads=8, dropout=0.1, batch_first=True)
        self.combiner_global = CombinerSimple(d, d * 4, d * 8, mode=mode)
        self.combiner_local = CombinerSimple(d, d * 4, d * 8, mode=mode)
        self.combiner = CombinerSimple(d, d * 4, d * 8, mode=mode)
        self.mode = mode
        self.max_batch = 2048                      # queries per kernel chain (bounds the workspace: ~1.5 MB per query;
                                                   # measured 14.3 / 13.7 / 13.2 / 13.1 ms per 4096 queries at 512 / 1024 / 2048 / 4096)
        self._packed: Optional[torch.Tensor] = None
        self._versions = None
        if device is not None:
            self.to(device)

    def invalidate_cache(self) -> None:
        """Drop the packed bf16 weight copies.  They are refreshed automatically when a parameter's version counter
        changes (``load_state_dict``, ``copy_``, optimizer steps); call this after writing through ``.data``."""
        self._packed = None
        self._versions = None

    def set_mode(self, mode: str) -> "DVR_module":
        if mode not in ("bf16", "fp32"):
            raise ErnError(f"unknown mode {mode!r}")
        self.mode = mode
        for m in (self.SR_module, self.combiner_global, self.combiner_local, self.combiner):
            m.set_mode(mode)
        return self

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts reference checkpoints: drops HF's ``position_ids`` buffers (persistent in transformers <= 4.30) and
        tolerates a missing ``cls_token`` (on CUDA the reference's ``nn.Parameter(...).to(device)`` is a plain tensor
        and never reaches the checkpoint)."""
        sd = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}
        if "transformer_layer.cls_token" not in sd:
            sd["transformer_layer.cls_token"] = self.transformer_layer.cls_token.detach()
        return super().load_state_dict(sd, strict=strict, **kw)

    # ---------------------------------------------------------------------------------------------
    def _weights(self, device) -> DvrWeights:
        bm = self.transformer_layer.bert_encoder.bert_model
        w = DvrWeights()
        tensors = [self.transformer_layer.cls_token, bm.embeddings.position_embeddings.weight,
                   bm.embeddings.token_type_embeddings.weight, bm.embeddings.LayerNorm.weight, bm.embeddings.LayerNorm.bias]
        w.cls_token, w.pos_emb, w.type_emb, w.emb_ln_w, w.emb_ln_b = [t.data_ptr() for t in tensors]
        w.n_layers, w.intermediate = self.n_layers, _INTERMEDIATE
        for i, layer in enumerate(bm.encoder.layer):
            a, o = layer.attention, layer.output
            sa = getattr(a, "self")
            ts = [sa.query.weight, sa.query.bias, sa.key.weight, sa.key.bias, sa.value.weight, sa.value.bias,
                  a.output.dense.weight, a.output.dense.bias, a.output.LayerNorm.weight, a.output.LayerNorm.bias,
                  layer.intermediate.dense.weight, layer.intermediate.dense.bias,
                  o.dense.weight, o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias]
            w.layers[i] = BertLayerWeights(*[t.data_ptr() for t in ts])
            tensors += ts
        mr = self.MR_component
        ts = [mr.in_proj_weight, mr.in_proj_bias, mr.out_proj.weight, mr.out_proj.bias]
        w.mha_in_w, w.mha_in_b, w.mha_out_w, w.mha_out_b = [t.data_ptr() for t in ts]
        tensors += ts
        for t in tensors:
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ErnError("DVR parameters must be contiguous float32 on the input's CUDA device")
        if self.mode == "bf16":
            versions = tuple((t.data_ptr(), t._version) for t in tensors)
            if self._packed is None or self._versions != versions or self._packed.device != device:
                nbytes = L.lib().ern_dvr_packed_bytes(self.dim, _INTERMEDIATE, self.n_layers)
                packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
                L.check(L.lib().ern_dvr_pack(C.byref(w), self.dim, packed.data_ptr(), L.stream_ptr(device)))
                launch_counter.add(6 * self.n_layers + 2)
                self._packed, self._versions = packed, versions
            w.packed_bf16 = self._packed.data_ptr()
        return w

    def encode(self, ref_patch_features: torch.Tensor, text_seq_features: torch.Tensor):
        """(patches [B,P,D], tokens [B,T,D]) -> (cross_vision_feats[:, :P] [B,P,D], seq_text_mean [B,D])
        -- models/fusion_model.py:35-49 without the SR/combiner tail."""
        L.require_cuda(ref_patch_features, "ref_patch_features")
        L.require_cuda(text_seq_features, "text_seq_features")
        x = ref_patch_features.detach().float().contiguous()
        t = text_seq_features.detach().float().contiguous()
        if x.dim() != 3 or t.dim() != 3 or x.shape[0] != t.shape[0] or x.shape[2] != self.dim or t.shape[2] != self.dim:
            raise ErnError(f"expected [B,P,{self.dim}] and [B,T,{self.dim}], got {tuple(x.shape)} and {tuple(t.shape)}")
        dev, B, P, T = x.device, x.shape[0], x.shape[1], t.shape[1]
        mode = MODE_BF16 if self.mode == "bf16" else MODE_FP32
        lib = L.lib()
        cross = torch.empty((B, P, self.dim), dtype=torch.float32, device=dev)
        seq_mean = torch.empty((B, self.dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            w = self._weights(dev)
            step = max(1, min(self.max_batch, B))
            wsb = lib.ern_dvr_workspace_bytes(step, P, T, self.dim, _INTERMEDIATE, mode)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
            for s in range(0, B, step):
                n = min(step, B - s)
                L.check(lib.ern_dvr_encode(C.byref(w), self.dim, self.heads, P, T, mode, x[s:s + n].data_ptr(),
                                           t[s:s + n].data_ptr(), n, cross[s:s + n].data_ptr(), seq_mean[s:s + n].data_ptr(),
                                           ws.data_ptr(), wsb, L.stream_ptr(dev)))
                launch_counter.add(1 + 11 * self.n_layers + 5)
        return cross, seq_mean

    def forward(self, ref_patch_features, text_seq_features, ref_global_feats, text_global_feats):
        """models/fusion_model.py:26-55 -> fused query features [B, D] (unit norm)."""
        if self.training:
            raise ErnError("DVR_module (B200) implements the eval-mode forward only: call model.eval()")
        cross, seq_text_mean = self.encode(ref_patch_features, text_seq_features)
        patch_vision_mean = self.SR_module(cross)                                      # :48
        global_feats = self.combiner_global(ref_global_feats, text_global_feats)       # :52
        local_feats = self.combiner_local(patch_vision_mean, seq_text_mean)            # :53
        return self.combiner(global_feats, local_feats)                                # :54
