"""Sharding policies for the falcon family.  Parity: reference `colossalai/shardformer/policies/falcon.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class FalconModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.falcon.FalconModel`."""


class FalconForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.falcon.FalconForCausalLM`."""


class FalconForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.falcon.FalconForSequenceClassification`."""


class FalconForTokenClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.falcon.FalconForTokenClassification`."""


class FalconForQuestionAnsweringPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.falcon.FalconForQuestionAnswering`."""


__all__ = ['FalconModelPolicy', 'FalconForCausalLMPolicy', 'FalconForSequenceClassificationPolicy', 'FalconForTokenClassificationPolicy', 'FalconForQuestionAnsweringPolicy']
