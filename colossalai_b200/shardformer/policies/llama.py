"""Sharding policies for the llama family.  Parity: reference `colossalai/shardformer/policies/llama.py:30-400`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class LlamaModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.llama.LlamaModel`."""


class LlamaForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.llama.LlamaForCausalLM`."""


class LlamaForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.llama.LlamaForSequenceClassification`."""


__all__ = ['LlamaModelPolicy', 'LlamaForCausalLMPolicy', 'LlamaForSequenceClassificationPolicy']
