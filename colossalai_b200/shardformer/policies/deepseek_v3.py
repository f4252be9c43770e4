"""Sharding policies for the deepseek_v3 family.  Parity: reference `colossalai/shardformer/policies/deepseek_v3.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class DeepseekV3ModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.deepseek_v3.DeepseekV3Model`."""


class DeepseekV3ForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.deepseek_v3.DeepseekV3ForCausalLM`."""


__all__ = ['DeepseekV3ModelPolicy', 'DeepseekV3ForCausalLMPolicy']
