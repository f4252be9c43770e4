"""Sharding policies for the chatglm family.  Parity: reference `colossalai/shardformer/policies/chatglm2.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class ChatGLMModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.chatglm.ChatGLMModel`."""


class ChatGLMForConditionalGenerationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.chatglm.ChatGLMForConditionalGeneration`."""


__all__ = ['ChatGLMModelPolicy', 'ChatGLMForConditionalGenerationPolicy']
