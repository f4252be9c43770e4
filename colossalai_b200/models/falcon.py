"""Falcon (parallel block, multi-query attention).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/falcon.py; modeling/falcon.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "falcon-7b"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class FalconModel(TransformerBackboneModel):
    """FalconModel — `TransformerBackboneModel` specialised for the falcon family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class FalconForCausalLM(TransformerLMHeadModel):
    """FalconForCausalLM — `TransformerLMHeadModel` specialised for the falcon family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class FalconForSequenceClassification(TransformerForSequenceClassification):
    """FalconForSequenceClassification — `TransformerForSequenceClassification` specialised for the falcon family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class FalconForTokenClassification(TransformerForTokenClassification):
    """FalconForTokenClassification — `TransformerForTokenClassification` specialised for the falcon family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class FalconForQuestionAnswering(TransformerForQuestionAnswering):
    """FalconForQuestionAnswering — `TransformerForQuestionAnswering` specialised for the falcon family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'FalconModel', 'FalconForCausalLM', 'FalconForSequenceClassification', 'FalconForTokenClassification', 'FalconForQuestionAnswering']
