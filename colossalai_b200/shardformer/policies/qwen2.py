"""Sharding policies for the qwen2 family.  Parity: reference `colossalai/shardformer/policies/qwen2.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class Qwen2ModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.qwen2.Qwen2Model`."""


class Qwen2ForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.qwen2.Qwen2ForCausalLM`."""


class Qwen2ForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.qwen2.Qwen2ForSequenceClassification`."""


__all__ = ['Qwen2ModelPolicy', 'Qwen2ForCausalLMPolicy', 'Qwen2ForSequenceClassificationPolicy']
