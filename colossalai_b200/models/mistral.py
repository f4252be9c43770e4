"""Mistral (Llama block + sliding-window attention).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/mistral.py; modeling/mistral.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "mistral-7b"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class MistralModel(TransformerBackboneModel):
    """MistralModel — `TransformerBackboneModel` specialised for the mistral family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class MistralForCausalLM(TransformerLMHeadModel):
    """MistralForCausalLM — `TransformerLMHeadModel` specialised for the mistral family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class MistralForSequenceClassification(TransformerForSequenceClassification):
    """MistralForSequenceClassification — `TransformerForSequenceClassification` specialised for the mistral family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'MistralModel', 'MistralForCausalLM', 'MistralForSequenceClassification']
