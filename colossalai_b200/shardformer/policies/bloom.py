"""Sharding policies for the bloom family.  Parity: reference `colossalai/shardformer/policies/bloom.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class BloomModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bloom.BloomModel`."""


class BloomForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bloom.BloomForCausalLM`."""


class BloomForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bloom.BloomForSequenceClassification`."""


class BloomForTokenClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bloom.BloomForTokenClassification`."""


class BloomForQuestionAnsweringPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bloom.BloomForQuestionAnswering`."""


__all__ = ['BloomModelPolicy', 'BloomForCausalLMPolicy', 'BloomForSequenceClassificationPolicy', 'BloomForTokenClassificationPolicy', 'BloomForQuestionAnsweringPolicy']
