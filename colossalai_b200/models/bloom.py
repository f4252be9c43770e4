"""BLOOM (ALiBi, LayerNorm, embedding LayerNorm).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/bloom.py; modeling/bloom.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "bloom-560m"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class BloomModel(TransformerBackboneModel):
    """BloomModel — `TransformerBackboneModel` specialised for the bloom family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BloomForCausalLM(TransformerLMHeadModel):
    """BloomForCausalLM — `TransformerLMHeadModel` specialised for the bloom family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BloomForSequenceClassification(TransformerForSequenceClassification):
    """BloomForSequenceClassification — `TransformerForSequenceClassification` specialised for the bloom family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BloomForTokenClassification(TransformerForTokenClassification):
    """BloomForTokenClassification — `TransformerForTokenClassification` specialised for the bloom family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BloomForQuestionAnswering(TransformerForQuestionAnswering):
    """BloomForQuestionAnswering — `TransformerForQuestionAnswering` specialised for the bloom family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'BloomModel', 'BloomForCausalLM', 'BloomForSequenceClassification', 'BloomForTokenClassification', 'BloomForQuestionAnswering']
