"""B200-native (sm_100a) implementation of FashionERN's composed-retrieval scoring path.

Public surface (mirrors the reference's own, see DESIGN.md / INTEGRATION.md):

  CombinerSimple, accelerate_ern            fusion head            (models/fusion_model.py:58-94)
  VisualSR                                  patch attention pooling (models/fusion_model.py:97-154; 'next' row)
  compute_{fiq,shoes,200k,cirr}_val_metrics metric tails           (run/test/test_*.py, run/valid/validate_*.py)
  compute_val_metrics, score_topk_recall
  BatchBasedClassificationLoss              training criterion      (losses/loss.py:6-14; 'next' row, fwd + bwd)
  ops.*                                     tensor-level wrappers of the C ABI (include/ern_b200.h)
  sharded.*                                 row-sharded gallery over the GPUs of one box

All compute is hand-written CUDA behind ``libern_b200.so``; there is no CPU or torch fallback.
"""
from ._lib import ErnError, MODE_BF16, MODE_FP32, RANK_REFERENCE, RANK_SIMILARITY  # noqa: F401
from .combiner import CombinerSimple, accelerate_ern  # noqa: F401
from .visual_sr import VisualSR  # noqa: F401
from .dvr import DVR_module  # noqa: F401
from .model import ERN  # noqa: F401
from .loss import BatchBasedClassificationLoss  # noqa: F401
from . import ops, sharded, store  # noqa: F401
from .metrics import (compute_200k_val_metrics, compute_cirr_val_metrics, compute_fiq_val_metrics,  # noqa: F401
                      compute_shoes_val_metrics, compute_val_metrics, score_topk_recall, set_precision,
                      set_tokenizer_factory)
