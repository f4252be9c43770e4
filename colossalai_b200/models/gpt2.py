"""GPT-2 (learned positions, LayerNorm, GELU MLP, tied embeddings).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/gpt2.py; modeling/gpt2.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "gpt2"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class GPT2Model(TransformerBackboneModel):
    """GPT2Model — `TransformerBackboneModel` specialised for the gpt2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPT2LMHeadModel(TransformerLMHeadModel):
    """GPT2LMHeadModel — `TransformerLMHeadModel` specialised for the gpt2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPT2DoubleHeadsModel(TransformerLMHeadModel):
    """GPT2DoubleHeadsModel — `TransformerLMHeadModel` specialised for the gpt2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPT2ForQuestionAnswering(TransformerForQuestionAnswering):
    """GPT2ForQuestionAnswering — `TransformerForQuestionAnswering` specialised for the gpt2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPT2ForTokenClassification(TransformerForTokenClassification):
    """GPT2ForTokenClassification — `TransformerForTokenClassification` specialised for the gpt2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPT2ForSequenceClassification(TransformerForSequenceClassification):
    """GPT2ForSequenceClassification — `TransformerForSequenceClassification` specialised for the gpt2 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'GPT2Model', 'GPT2LMHeadModel', 'GPT2DoubleHeadsModel', 'GPT2ForQuestionAnswering', 'GPT2ForTokenClassification', 'GPT2ForSequenceClassification']
