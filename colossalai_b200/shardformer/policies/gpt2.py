"""Sharding policies for the gpt2 family.  Parity: reference `colossalai/shardformer/policies/gpt2.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class GPT2ModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gpt2.GPT2Model`."""


class GPT2LMHeadModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gpt2.GPT2LMHeadModel`."""


class GPT2DoubleHeadsModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gpt2.GPT2DoubleHeadsModel`."""


class GPT2ForQuestionAnsweringPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gpt2.GPT2ForQuestionAnswering`."""


class GPT2ForTokenClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gpt2.GPT2ForTokenClassification`."""


class GPT2ForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.gpt2.GPT2ForSequenceClassification`."""


__all__ = ['GPT2ModelPolicy', 'GPT2LMHeadModelPolicy', 'GPT2DoubleHeadsModelPolicy', 'GPT2ForQuestionAnsweringPolicy', 'GPT2ForTokenClassificationPolicy', 'GPT2ForSequenceClassificationPolicy']
