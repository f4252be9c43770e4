"""Llama / Llama-2 / Llama-3 (RMSNorm, RoPE, SwiGLU, GQA).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/llama.py:30-400; modeling/llama.py:43-600`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "llama2-7b"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class LlamaModel(TransformerBackboneModel):
    """LlamaModel — `TransformerBackboneModel` specialised for the llama family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class LlamaForCausalLM(TransformerLMHeadModel):
    """LlamaForCausalLM — `TransformerLMHeadModel` specialised for the llama family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class LlamaForSequenceClassification(TransformerForSequenceClassification):
    """LlamaForSequenceClassification — `TransformerForSequenceClassification` specialised for the llama family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'LlamaModel', 'LlamaForCausalLM', 'LlamaForSequenceClassification']
