"""Qwen2 (Llama block with QKV bias).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/qwen2.py; modeling/qwen2.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "qwen2-7b"
FAMILY_DEFAULTS = {'attention_bias': True, 'attention_out_bias': False}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class Qwen2Model(TransformerBackboneModel):
    """Qwen2Model — `TransformerBackboneModel` specialised for the qwen2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class Qwen2ForCausalLM(TransformerLMHeadModel):
    """Qwen2ForCausalLM — `TransformerLMHeadModel` specialised for the qwen2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class Qwen2ForSequenceClassification(TransformerForSequenceClassification):
    """Qwen2ForSequenceClassification — `TransformerForSequenceClassification` specialised for the qwen2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'Qwen2Model', 'Qwen2ForCausalLM', 'Qwen2ForSequenceClassification']
