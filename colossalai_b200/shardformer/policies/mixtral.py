"""Sharding policies for the mixtral family.  Parity: reference `colossalai/shardformer/policies/mixtral.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class MixtralModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.mixtral.MixtralModel`."""


class MixtralForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.mixtral.MixtralForCausalLM`."""


__all__ = ['MixtralModelPolicy', 'MixtralForCausalLMPolicy']
