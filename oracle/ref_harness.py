"""Run the UNMODIFIED reference hot-path functions on synthetic features.  TEST INFRASTRUCTURE ONLY.

Works only where ``/root/reference`` is mounted (the authoring container).  It never runs on the
GPU box; what it produces travels as fixtures under ``tests/golden/`` (see ``make_golden.py``).

How: ``open_clip`` is not installable offline, so a stub module that only offers
``get_tokenizer`` is placed in ``sys.modules``; datasets and the CLIP text tower are replaced by
fakes that serve seeded synthetic tensors with the tuple layouts of the reference's data loaders:

  FashionIQ  (ref_name, target_name, [cap1, cap2], ref_patch[13,D])                dataloader/fashioniq.py:86
  Shoes      (ref_name, target_name, caption, ref_patch, tar_patch)                 dataloader/shoes.py:41
  CIRR       (ref_name, target_name, caption, ref_patch, [6 member names])          dataloader/cirr.py:73
  Fashion200k(ref_img, ref_id, modifier, targ_id, len(modifier), ref_patch)         dataloader/fashion200k_patch.py:354

Everything downstream -- ``generate_*_val_predictions``, ``ERN`` (BERT, MHA, VisualSR, the four
``CombinerSimple``), ``compute_*_val_metrics`` -- is the reference's own code, called verbatim.
"""
from __future__ import annotations

import contextlib
import re
import sys
import types
import warnings
from typing import Dict, List

import torch

REFERENCE_ROOT = "/root/reference"


def install() -> None:
    if "open_clip" not in sys.modules:
        stub = types.ModuleType("open_clip")

        def get_tokenizer(_name):
            def tok(texts, context_length=77):
                if isinstance(texts, str):
                    texts = [texts]
                out = torch.zeros(len(texts), context_length, dtype=torch.long)
                for i, t in enumerate(texts):
                    m = re.search(r"[qQ](\d+)", t)
                    out[i, 0] = int(m.group(1))
                return out
            return tok

        stub.get_tokenizer = get_tokenizer
        stub.create_model_and_transforms = None
        sys.modules["open_clip"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore")


class FakeClip:
    """Stands in for the unreleased TME CLIP fork: ``encode_text(tokens, mode=, visual_emb=)``
    as called at run/test/test_fiq.py:102-103."""

    def __init__(self, text_global: torch.Tensor, text_seq: torch.Tensor):
        self.text_global, self.text_seq = text_global, text_seq

    def eval(self):
        return self

    def encode_text(self, tokens, mode="global", visual_emb=None):
        idx = tokens[:, 0].long().cpu()
        if mode == "seq":
            return self.text_seq[idx]
        return self.text_global[idx], None


class FakeRelative(torch.utils.data.Dataset):
    def __init__(self, kind: str, ref_names, target_names, ref_patches, group_members=None):
        self.kind, self.ref, self.tgt, self.patch, self.members = kind, ref_names, target_names, ref_patches, group_members

    def __len__(self):
        return len(self.ref)

    def __getitem__(self, i):
        if self.kind == "fiq":
            return self.ref[i], self.tgt[i], [f"q{i}.", "x"], self.patch[i]
        if self.kind == "shoes":
            return self.ref[i], self.tgt[i], f"q{i}", self.patch[i], self.patch[i]
        if self.kind == "cirr":
            return self.ref[i], self.tgt[i], f"q{i}", self.patch[i], list(self.members[i])
        if self.kind == "200k":
            return 0, self.ref[i], f"q{i}", self.tgt[i], 2, self.patch[i]
        raise ValueError(self.kind)


class Recorder:
    """Wraps an ``ERN`` and records what crosses the hot-path boundary."""

    def __init__(self, model):
        self.model = model
        self.pred: List[torch.Tensor] = []
        self.index_in = None
        self.index_local = None
        self.index_out = None
        self.sr_out = None
        self.combiner_io: Dict[str, List] = {}

    def hook_combiners(self):
        def mk(name):
            def hook(_m, inp, out):
                self.combiner_io.setdefault(name, []).append((inp[0].detach().clone(), inp[1].detach().clone(), out.detach().clone()))
            return hook
        for name, mod in (("DVR.combiner_global", self.model.DVR.combiner_global),
                          ("DVR.combiner_local", self.model.DVR.combiner_local),
                          ("DVR.combiner", self.model.DVR.combiner),
                          ("Combiner_module", self.model.Combiner_module)):
            mod.register_forward_hook(mk(name))
        self.model.SR_module.register_forward_hook(lambda _m, _i, out: setattr(self, "sr_out", out.detach().clone()))

    def __call__(self, **kw):
        out = self.model(**kw)
        if kw.get("mode") == "test":
            self.pred.append(out.detach().clone())
        elif kw.get("mode") == "index":
            self.index_in, self.index_local, self.index_out = kw["tar_feats"], kw["tar_local_feats"], out.detach().clone()
        return out


@contextlib.contextmanager
def capture_argsort(store: list):
    """Record the result of the reference's ``torch.argsort`` call (run/test/test_fiq.py:50)."""
    orig = torch.argsort

    def spy(*a, **k):
        r = orig(*a, **k)
        store.append(r.detach().clone())
        return r

    torch.argsort = spy
    try:
        yield
    finally:
        torch.argsort = orig


def build_ern(dim: int, combiner_states: Dict[str, Dict[str, torch.Tensor]], seed: int = 0):
    """Reference ``ERN`` on CPU in eval/fp32 (run/test/test_fiq.py:148,168-169) with the four combiners'
    parameters overwritten by the given synthetic state dicts."""
    install()
    from models.model import ERN
    torch.manual_seed(seed)
    model = ERN(FakeClip(None, None), dim, "cpu")
    for name, sd in combiner_states.items():
        mod = model
        for part in name.split("."):
            mod = getattr(mod, part)
        mod.load_state_dict(sd)
    model.eval()
    return model.float()


def reference_combiner(dim: int, sd: Dict[str, torch.Tensor]):
    install()
    from models.fusion_model import CombinerSimple
    m = CombinerSimple(dim, dim * 4, dim * 8)
    m.load_state_dict(sd)
    return m.eval().float()


def metric_fn(kind: str, variant: str = "test"):
    """The reference function object for a dataset kind ('fiq','shoes','200k','cirr','val')."""
    install()
    import importlib
    if variant == "test":
        mod = {"fiq": "run.test.test_fiq", "shoes": "run.test.test_shoes", "200k": "run.test.test_200k",
               "cirr": "run.test.test_cirr", "val": "run.test.test_val"}[kind]
        fn = {"fiq": "compute_fiq_val_metrics", "shoes": "compute_shoes_val_metrics",
              "200k": "compute_200k_val_metrics", "cirr": "compute_cirr_val_metrics",
              "val": "compute_fiq_val_metrics"}[kind]
    else:
        mod = {"fiq": "run.valid.validate_fiq", "shoes": "run.valid.validate_shoes",
               "cirr": "run.valid.validate_cirr"}[kind]
        fn = {"fiq": "compute_fiq_val_metrics", "shoes": "compute_shoes_val_metrics",
              "cirr": "compute_cirr_val_metrics"}[kind]
    with contextlib.redirect_stdout(None):
        m = importlib.import_module(mod)
    return getattr(m, fn)


def run_metric(kind: str, dataset, clip, index_features, index_local, index_names, model, dim: int,
               batch_size: int = 32):
    """Call the reference's 11-argument ``compute_*_val_metrics`` verbatim on CPU."""
    fn = metric_fn(kind)
    with contextlib.redirect_stdout(None), torch.no_grad():
        return fn(dataset, clip, index_features, index_local, index_names, model, "cpu", dim,
                  batch_size, 0, "RN50x4")
