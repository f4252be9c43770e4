"""Sharding policies for the mistral family.  Parity: reference `colossalai/shardformer/policies/mistral.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class MistralModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.mistral.MistralModel`."""


class MistralForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.mistral.MistralForCausalLM`."""


class MistralForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.mistral.MistralForSequenceClassification`."""


__all__ = ['MistralModelPolicy', 'MistralForCausalLMPolicy', 'MistralForSequenceClassificationPolicy']
