"""Drop-in replacement of the reference's ``ERN`` wrapper (models/model.py:7-75) for evaluation.

Same constructor ``ERN(clip_model, feature_dim, device)``, same attribute names (``DVR``, ``SR_module``,
``Combiner_module``; ``image_clip`` / ``text_clip`` are thin pass-throughs to the caller's CLIP model, which is
outside this repository's scope) and therefore the same ``state_dict`` keys, same ``forward(..., mode=...)`` dispatch.
``mode="index"`` (gallery side) and ``mode="test"`` (query side) run entirely on the B200 kernels; the training
branch is refused (no backward on the accelerated path).
"""
from __future__ import annotations

import warnings

import torch
from torch import nn

from ._lib import ErnError
from .combiner import CombinerSimple
from .dvr import DVR_module
from .visual_sr import VisualSR


class _ClipPassThrough(nn.Module):
    """models/clip_model.py:5-31: call the caller's CLIP towers under no_grad; holds no parameters of its own."""

    def __init__(self, clip_model, kind: str):
        super().__init__()
        object.__setattr__(self, "_clip", clip_model)     # not registered: the backbone is not part of this state_dict
        self.kind = kind

    def forward(self, x, mode="global", visual_emb=None):
        clip = object.__getattribute__(self, "_clip")
        if hasattr(clip, "eval"):
            clip.eval()                                   # models/clip_model.py:11,24
        with torch.no_grad():
            if self.kind == "image":
                return clip.encode_image(x)
            if mode == "seq":
                return clip.encode_text(x, mode="seq", visual_emb=visual_emb)
            return clip.encode_text(x, visual_emb=visual_emb)


class ERN(nn.Module):
    def __init__(self, clip_model, feature_dim, device, mode: str = "bf16"):
        super().__init__()
        self.image_clip = _ClipPassThrough(clip_model, "image")
        self.text_clip = _ClipPassThrough(clip_model, "text")
        self.DVR = DVR_module(feature_dim=feature_dim, device=None, mode=mode)
        self.SR_module = VisualSR(embed_dim=feature_dim, mode=mode)
        self.Combiner_module = CombinerSimple(feature_dim, feature_dim * 4, feature_dim * 8, mode=mode)
        if device is not None:
            self.to(device)

    def set_mode(self, mode: str) -> "ERN":
        self.DVR.set_mode(mode)
        self.SR_module.set_mode(mode)
        self.Combiner_module.set_mode(mode)
        return self

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """Same keys as the reference's: when the wrapped CLIP model is an ``nn.Module`` its weights appear under
        ``image_clip.clip_model.*`` and ``text_clip.clip_model.*`` (models/clip_model.py:8,21 register it twice)."""
        sd = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        clip = object.__getattribute__(self.image_clip, "_clip")
        if isinstance(clip, nn.Module):
            for tower in ("image_clip", "text_clip"):
                clip.state_dict(destination=sd, prefix=f"{prefix}{tower}.clip_model.", keep_vars=keep_vars)
        return sd

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Reference checkpoints (``torch.save(model.module.state_dict())``, run/train/train_fiq.py:174-175) may
        carry CLIP keys under ``image_clip.`` / ``text_clip.`` and HF ``position_ids`` buffers, and lack
        ``DVR.transformer_layer.cls_token`` when they were built on CUDA: all handled here."""
        clip = object.__getattribute__(self.image_clip, "_clip")
        clip_sd = {k.split("clip_model.", 1)[1]: v for k, v in state_dict.items()
                   if k.startswith(("image_clip.clip_model.", "text_clip.clip_model."))}
        if clip_sd:
            # the reference registers clip_model as a submodule (models/clip_model.py:8,19), so its checkpoints carry
            # the backbone and its load_state_dict restores it: forward those keys into the caller's CLIP model
            if isinstance(clip, nn.Module):
                missing, unexpected = clip.load_state_dict(clip_sd, strict=False)
                if unexpected or len(missing) == len(clip.state_dict()):
                    warnings.warn(f"checkpoint CLIP weights did not match the wrapped CLIP model "
                                  f"({len(unexpected)} unexpected, {len(missing)} missing keys)")
            else:
                warnings.warn("the checkpoint carries CLIP backbone weights (image_clip.* / text_clip.*) but the wrapped "
                              "clip_model is not an nn.Module: they were NOT loaded; make sure --clip-path matches")
        sd = {k: v for k, v in state_dict.items()
              if not k.startswith(("image_clip.", "text_clip.")) and not k.endswith("position_ids")}
        if "DVR.transformer_layer.cls_token" not in sd:
            sd["DVR.transformer_layer.cls_token"] = self.DVR.transformer_layer.cls_token.detach()
        return super().load_state_dict(sd, strict=strict, **kw)

    def forward(self, image=None, text=None, ref_feats=None, ref_local_feats=None, text_feats=None,
                text_seq_feats=None, tar_feats=None, tar_local_feats=None, mode="train"):
        if mode == "image":
            return self.image_clip(image)
        if mode == "text_global":
            return self.text_clip(text, mode="global", visual_emb=ref_local_feats)[0]
        if mode == "text_seq":
            return self.text_clip(text, mode="seq", visual_emb=ref_local_feats)
        if mode == "index":                                   # models/model.py:64-66
            return self.Combiner_module(tar_feats, self.SR_module(tar_local_feats))
        if mode == "test":                                    # models/model.py:68-69
            return self.DVR(ref_local_feats, text_seq_feats, ref_feats, text_feats)
        raise ErnError("ERN (B200) implements the evaluation modes only ('image', 'text_global', 'text_seq', 'index', "
                       "'test'); the training forward (models/model.py:71-75) is outside the accelerated path")
