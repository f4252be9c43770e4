"""Sharding policies for the command family.  Parity: reference `colossalai/shardformer/policies/command.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class CohereModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.command.CohereModel`."""


class CohereForCausalLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.command.CohereForCausalLM`."""


__all__ = ['CohereModelPolicy', 'CohereForCausalLMPolicy']
