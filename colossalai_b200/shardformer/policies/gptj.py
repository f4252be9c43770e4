"""Sharding policies for the gptj family.  Parity: reference `colossalai/shardformer/policies/gptj.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class GPTJModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gptj.GPTJModel`."""


class GPTJForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gptj.GPTJForCausalLM`."""


class GPTJForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gptj.GPTJForSequenceClassification`."""


class GPTJForQuestionAnsweringPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gptj.GPTJForQuestionAnswering`."""


__all__ = ['GPTJModelPolicy', 'GPTJForCausalLMPolicy', 'GPTJForSequenceClassificationPolicy', 'GPTJForQuestionAnsweringPolicy']
