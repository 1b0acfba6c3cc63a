"""Seeded synthetic inputs for the composed-retrieval scoring path.

There is no dataset or checkpoint access (no network), so every parity test, golden
fixture and benchmark feeds the path with features drawn from seeded torch CPU generators.
The same functions are used by ``oracle/make_golden.py`` (which runs the unmodified
reference on them) and by ``tests/`` / ``bench.py`` (which run the CUDA path), so both
sides see bit-identical inputs.  Shapes follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Tuple

import numpy as np
import torch

PATCHES = 13   # reference: run/test/test_fiq.py:130 (--patch-num)
TOKENS = 77    # reference: run/test/test_fiq.py:98  (context_length=77)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def combiner_state(seed: int, dim: int, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """State dict of one ``CombinerSimple(dim, 4*dim, 8*dim)`` with the reference's key names
    (models/fusion_model.py:73-84), values drawn like torch's default Linear init
    (uniform in +-1/sqrt(fan_in)) from an explicit generator so the draw does not depend on
    module construction order."""
    g = _gen(seed)
    proj, hid = 4 * dim, 8 * dim

    def lin(out_f, in_f):
        b = scale / np.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        bias = (torch.rand(out_f, generator=g) * 2 - 1) * b
        return w, bias

    sd: Dict[str, torch.Tensor] = {}
    sd["dynamic_scalar.0.weight"], sd["dynamic_scalar.0.bias"] = lin(hid, 2 * proj)
    sd["dynamic_scalar.3.weight"], sd["dynamic_scalar.3.bias"] = lin(1, hid)
    sd["text_projection_layer.0.weight"], sd["text_projection_layer.0.bias"] = lin(proj, dim)
    sd["image_projection_layer.0.weight"], sd["image_projection_layer.0.bias"] = lin(proj, dim)
    # make the gate informative: default init gives |logit| << 1, i.e. s ~ 0.5 everywhere
    sd["dynamic_scalar.3.weight"] = sd["dynamic_scalar.3.weight"] * 10.0
    return sd


def visualsr_state(seed: int, dim: int, patches: int = PATCHES) -> Dict[str, torch.Tensor]:
    """State dict of one ``VisualSR(dim)`` with the reference's key names (models/fusion_model.py:106-124):
    Xavier-uniform Linear weights like its ``init_weights`` (:126-134) but non-trivial biases and BatchNorm
    affine/running statistics, so that every term of the eval-mode forward is exercised."""
    g = _gen(seed)

    def xavier(out_f, in_f):
        r = float(np.sqrt(6.0) / np.sqrt(in_f + out_f))
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * r

    def bn(prefix, n, sd):
        sd[f"{prefix}.weight"] = 0.5 + torch.rand(n, generator=g)
        sd[f"{prefix}.bias"] = 0.2 * torch.randn(n, generator=g)
        sd[f"{prefix}.running_mean"] = 0.1 * torch.randn(n, generator=g)
        sd[f"{prefix}.running_var"] = 0.5 + torch.rand(n, generator=g)
        sd[f"{prefix}.num_batches_tracked"] = torch.tensor(100)

    sd: Dict[str, torch.Tensor] = {}
    sd["embedding_local.0.weight"] = xavier(dim, dim)
    sd["embedding_local.0.bias"] = 0.05 * torch.randn(dim, generator=g)
    bn("embedding_local.1", patches, sd)
    sd["embedding_global.0.weight"] = xavier(dim, dim)
    sd["embedding_global.0.bias"] = 0.05 * torch.randn(dim, generator=g)
    bn("embedding_global.1", dim, sd)
    sd["embedding_common.weight"] = xavier(1, dim) * 4.0
    sd["embedding_common.bias"] = 0.1 * torch.randn(1, generator=g)
    return sd


BERT_LAYERS = 2            # models/fusion_model.py:14  (PlusModel(..., layers=2))
BERT_INTERMEDIATE = 3072   # BertConfig default intermediate_size (models/fusion_model.py:162-170 does not set it)
BERT_PREFIX = "transformer_layer.bert_encoder.bert_model."


def dvr_state(seed: int, dim: int) -> Dict[str, torch.Tensor]:
    """State dict of the transformer / cross-attention part of one ``DVR_module(dim)`` with the reference's key
    names (models/fusion_model.py:8-24,157-216: HF ``BertModel`` without word embeddings, ``cls_token``,
    ``nn.MultiheadAttention`` ``MR_component``).  Values are drawn from an explicit generator at BERT-like scales
    (weights N(0, 0.02^2) scaled up a little so that attention is not uniform, LayerNorm affine near identity)."""
    g = _gen(seed)

    def normal(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    sd: Dict[str, torch.Tensor] = {}
    sd["transformer_layer.cls_token"] = normal(1, 1, dim, std=0.5)
    e = BERT_PREFIX + "embeddings."
    sd[e + "position_embeddings.weight"] = normal(512, dim, std=0.1)
    sd[e + "token_type_embeddings.weight"] = normal(2, dim, std=0.1)
    sd[e + "LayerNorm.weight"] = 1.0 + normal(dim, std=0.1)
    sd[e + "LayerNorm.bias"] = normal(dim, std=0.1)
    for l in range(BERT_LAYERS):
        p = f"{BERT_PREFIX}encoder.layer.{l}."
        for nm in ("query", "key", "value"):
            sd[p + f"attention.self.{nm}.weight"] = normal(dim, dim, std=0.06)
            sd[p + f"attention.self.{nm}.bias"] = normal(dim, std=0.05)
        sd[p + "attention.output.dense.weight"] = normal(dim, dim, std=0.04)
        sd[p + "attention.output.dense.bias"] = normal(dim, std=0.05)
        sd[p + "attention.output.LayerNorm.weight"] = 1.0 + normal(dim, std=0.1)
        sd[p + "attention.output.LayerNorm.bias"] = normal(dim, std=0.1)
        sd[p + "intermediate.dense.weight"] = normal(BERT_INTERMEDIATE, dim, std=0.04)
        sd[p + "intermediate.dense.bias"] = normal(BERT_INTERMEDIATE, std=0.05)
        sd[p + "output.dense.weight"] = normal(dim, BERT_INTERMEDIATE, std=0.03)
        sd[p + "output.dense.bias"] = normal(dim, std=0.05)
        sd[p + "output.LayerNorm.weight"] = 1.0 + normal(dim, std=0.1)
        sd[p + "output.LayerNorm.bias"] = normal(dim, std=0.1)
    sd[BERT_PREFIX + "pooler.dense.weight"] = normal(dim, dim)
    sd[BERT_PREFIX + "pooler.dense.bias"] = normal(dim)
    sd["MR_component.in_proj_weight"] = normal(3 * dim, dim, std=0.08)
    sd["MR_component.in_proj_bias"] = normal(3 * dim, std=0.05)
    sd["MR_component.out_proj.weight"] = normal(dim, dim, std=0.05)
    sd["MR_component.out_proj.bias"] = normal(dim, std=0.05)
    return sd


def dvr_full_state(seed: int, dim: int) -> Dict[str, torch.Tensor]:
    """Complete ``DVR_module`` state dict: transformer + MR_component + SR_module + the three fusion heads."""
    sd = dvr_state(seed, dim)
    sd.update({f"SR_module.{k}": v for k, v in visualsr_state(seed + 1, dim).items()})
    for i, name in enumerate(("combiner_global", "combiner_local", "combiner")):
        sd.update({f"{name}.{k}": v for k, v in combiner_state(seed + 2 + i, dim).items()})
    return sd


def ern_full_state(seed: int, dim: int) -> Dict[str, torch.Tensor]:
    """Complete ``ERN`` state dict (models/model.py:16-20) minus the CLIP backbone: DVR.*, SR_module.*, Combiner_module.*"""
    sd = {f"DVR.{k}": v for k, v in dvr_full_state(seed, dim).items()}
    sd.update({f"SR_module.{k}": v for k, v in visualsr_state(seed + 20, dim).items()})
    sd.update({f"Combiner_module.{k}": v for k, v in combiner_state(seed + 21, dim).items()})
    return sd


def features(seed: int, rows: int, dim: int, unit: bool = False) -> torch.Tensor:
    x = torch.randn(rows, dim, generator=_gen(seed))
    if unit:
        x = torch.nn.functional.normalize(x, dim=-1)
    return x


def loss_pair(seed: int, rows: int, dim: int, spread: float = 0.25, noise: float = 0.25
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(predicted, target) unit-norm features of one training batch.  Targets share a common direction (pairwise
    cosine ~0.94, like CLIP embeddings of one product category) and row i of ``predicted`` is a noisy copy of target
    i (cosine ~0.97), so the 100x-scaled in-batch classification loss (losses/loss.py:10-14) is neither saturated
    nor at chance."""
    common = features(seed + 2, 1, dim, unit=True)
    tar = torch.nn.functional.normalize(common + spread * features(seed, rows, dim, unit=True), dim=-1)
    pred = torch.nn.functional.normalize(tar + noise * features(seed + 1, rows, dim, unit=True), dim=-1)
    return pred, tar


def patch_features(seed: int, rows: int, dim: int, patches: int = PATCHES) -> torch.Tensor:
    return torch.randn(rows, patches, dim, generator=_gen(seed))


def token_features(seed: int, rows: int, dim: int, tokens: int = TOKENS) -> torch.Tensor:
    return torch.randn(rows, tokens, dim, generator=_gen(seed))


def unique_names(n: int, fmt: str = "B{:07d}") -> List[str]:
    return [fmt.format(i) for i in range(n)]


def caption_names(seed: int, n: int, classes: int) -> List[str]:
    """Fashion200k-style non-unique gallery 'names' (captions; dataloader/fashion200k_patch.py:287)."""
    cls = torch.randint(0, classes, (n,), generator=_gen(seed)).tolist()
    return [f"caption {c}" for c in cls]


def planted_ranks(seed: int, q: int, max_rank: int = 100) -> torch.Tensor:
    """Target rank per query with forced mass at the K boundaries."""
    g = _gen(seed)
    r = torch.randint(0, max_rank, (q,), generator=g)
    edges = torch.tensor([0, 4, 5, 9, 10, 49, 50])
    pick = torch.rand(q, generator=g) < 0.5
    r[pick] = edges[torch.randint(0, len(edges), (int(pick.sum()),), generator=g)]
    return r


def tensor_digest(t: torch.Tensor) -> str:
    """Stable fingerprint used to check that a regenerated tensor equals the one a golden was made from."""
    a = t.detach().cpu().contiguous().numpy()
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]
