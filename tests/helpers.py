"""Shared test helpers: golden loading and stand-ins for the off-path parts of the reference model."""
import json
import os

import numpy as np
import torch

from fashionern_aaai2024_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COMBINERS = ("DVR.combiner_global", "DVR.combiner_local", "DVR.combiner", "Combiner_module")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def case_inputs(meta):
    """Regenerate the seeded inputs a golden was made from (oracle/make_golden.py:case_inputs)."""
    kind, dim, q, n, seed = meta["kind"], meta["dim"], meta["q"], meta["n"], meta["seed"]
    index_features = syn.features(seed + 1, n, dim)
    index_local = syn.patch_features(seed + 2, n, dim)
    assert syn.tensor_digest(index_features) == meta["digest_index_features"], "torch RNG stream changed"
    assert syn.tensor_digest(index_local) == meta["digest_index_local"], "torch RNG stream changed"
    if kind == "200k":
        names = syn.caption_names(seed + 6, n, classes=max(8, n // 6))
    elif kind == "cirr":
        names = syn.unique_names(n, "dev-{}-img")
    else:
        names = syn.unique_names(n)
    states = {name: syn.combiner_state(seed + 10 + i, dim) for i, name in enumerate(COMBINERS)}
    return index_features, index_local, names, states


class FakeClip:
    """encode_text stand-in (run/test/test_fiq.py:102-103): the query index rides in feature column 0."""

    def __init__(self, dim):
        self.dim = dim

    def encode_text(self, tokens, mode="global", visual_emb=None):
        idx = tokens[:, 0].float()
        if mode == "seq":
            out = torch.zeros(tokens.shape[0], 77, self.dim, device=tokens.device)
            out[:, 0, 0] = idx
            return out
        out = torch.zeros(tokens.shape[0], self.dim, device=tokens.device)
        out[:, 0] = idx
        return out, None


def fake_tokenizer_factory(_name):
    import re

    def tok(texts, context_length=77):
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            out[i, 0] = int(re.search(r"[qQ](\d+)", t).group(1))
        return out
    return tok


class FakeRelative(torch.utils.data.Dataset):
    """Same tuple layouts as the reference's data loaders (see oracle/ref_harness.py)."""

    def __init__(self, kind, ref_names, target_names, ref_patches, members=None):
        self.kind, self.ref, self.tgt, self.patch, self.members = kind, ref_names, target_names, ref_patches, members

    def __len__(self):
        return len(self.ref)

    def __getitem__(self, i):
        if self.kind == "fiq":
            return self.ref[i], self.tgt[i], [f"q{i}.", "x"], self.patch[i]
        if self.kind == "shoes":
            return self.ref[i], self.tgt[i], f"q{i}", self.patch[i], self.patch[i]
        if self.kind == "cirr":
            return self.ref[i], self.tgt[i], f"q{i}", self.patch[i], list(self.members[i])
        return 0, self.ref[i], f"q{i}", self.tgt[i], 2, self.patch[i]


class StandInERN(torch.nn.Module):
    """Mode dispatch of models/model.py:22-75 with the OFF-PATH parts served from the golden:
    mode="test" (DVR: BERT + MHA + ...) returns the recorded query features, the VisualSR output of
    mode="index" is the recorded one; the ON-PATH gallery-side fusion head is the real B200 module."""

    def __init__(self, combiner, pred, sr_out):
        super().__init__()
        self.Combiner_module = combiner
        self.register_buffer("pred", pred)
        self.register_buffer("sr_out", sr_out)

    def forward(self, ref_feats=None, ref_local_feats=None, text_feats=None, text_seq_feats=None,
                tar_feats=None, tar_local_feats=None, mode="train"):
        if mode == "test":
            return self.pred[text_feats[:, 0].long()]
        if mode == "index":
            return self.Combiner_module(tar_feats, self.sr_out)
        raise ValueError(mode)


class SeededClip:
    """encode_text stand-in serving seeded text features by query index (same as oracle/ref_harness.FakeClip)."""

    def __init__(self, text_global, text_seq):
        self.text_global, self.text_seq = text_global, text_seq

    def encode_text(self, tokens, mode="global", visual_emb=None):
        idx = tokens[:, 0].long().to(self.text_global.device)
        if mode == "seq":
            return self.text_seq[idx]
        return self.text_global[idx], None
