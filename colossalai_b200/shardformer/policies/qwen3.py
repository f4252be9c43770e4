"""Sharding policies for the qwen3 family.  Parity: reference `colossalai/shardformer/policies/qwen3.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class Qwen3ModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.qwen3.Qwen3Model`."""


class Qwen3ForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.qwen3.Qwen3ForCausalLM`."""


class Qwen3ForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.qwen3.Qwen3ForSequenceClassification`."""


__all__ = ['Qwen3ModelPolicy', 'Qwen3ForCausalLMPolicy', 'Qwen3ForSequenceClassificationPolicy']
