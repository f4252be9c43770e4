"""Sharding policies for the opt family.  Parity: reference `colossalai/shardformer/policies/opt.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class OPTModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.opt.OPTModel`."""


class OPTForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.opt.OPTForCausalLM`."""


class OPTForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.opt.OPTForSequenceClassification`."""


class OPTForQuestionAnsweringPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.opt.OPTForQuestionAnswering`."""


__all__ = ['OPTModelPolicy', 'OPTForCausalLMPolicy', 'OPTForSequenceClassificationPolicy', 'OPTForQuestionAnsweringPolicy']
