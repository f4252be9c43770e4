"""OPT (learned positions, LayerNorm, ReLU MLP).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/opt.py; modeling/opt.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "opt-125m"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class OPTModel(TransformerBackboneModel):
    """OPTModel — `TransformerBackboneModel` specialised for the opt family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class OPTForCausalLM(TransformerLMHeadModel):
    """OPTForCausalLM — `TransformerLMHeadModel` specialised for the opt family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class OPTForSequenceClassification(TransformerForSequenceClassification):
    """OPTForSequenceClassification — `TransformerForSequenceClassification` specialised for the opt family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class OPTForQuestionAnswering(TransformerForQuestionAnswering):
    """OPTForQuestionAnswering — `TransformerForQuestionAnswering` specialised for the opt family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'OPTModel', 'OPTForCausalLM', 'OPTForSequenceClassification', 'OPTForQuestionAnswering']
