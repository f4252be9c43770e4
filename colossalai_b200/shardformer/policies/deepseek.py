"""Sharding policies for the deepseek family.  Parity: reference `colossalai/shardformer/policies/deepseek.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class DeepseekModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.deepseek.DeepseekModel`."""


class DeepseekForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.deepseek.DeepseekForCausalLM`."""


__all__ = ['DeepseekModelPolicy', 'DeepseekForCausalLMPolicy']
