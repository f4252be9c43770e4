"""Sharding policies for the Baichuan family (fused `W_pack` QKV -> one fused column-parallel linear).
Parity: reference `colossalai/inference/modeling/policy/nopadding_baichuan.py:1-110`."""
from __future__ import annotations

from .transformer import TransformerPolicy

__all__ = ["BaichuanModelPolicy", "BaichuanForCausalLMPolicy", "BaichuanForSequenceClassificationPolicy"]


class BaichuanModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.baichuan.BaichuanModel`."""


class BaichuanForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.baichuan.BaichuanForCausalLM`."""

    def preprocess(self):
        # NormHead rows are normalised per vocabulary entry, which commutes with vocab-parallel sharding: fold first
        if getattr(self.model.cfg, "norm_head", False) and not self.model.training:
            self.model.fold_norm_head()
        return self.model


class BaichuanForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.baichuan.BaichuanForSequenceClassification`."""
