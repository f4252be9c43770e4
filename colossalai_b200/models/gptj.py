"""GPT-J (parallel attention+MLP block, interleaved partial RoPE).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/gptj.py; modeling/gptj.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "gptj-6b"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class GPTJModel(TransformerBackboneModel):
    """GPTJModel — `TransformerBackboneModel` specialised for the gptj family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPTJForCausalLM(TransformerLMHeadModel):
    """GPTJForCausalLM — `TransformerLMHeadModel` specialised for the gptj family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPTJForSequenceClassification(TransformerForSequenceClassification):
    """GPTJForSequenceClassification — `TransformerForSequenceClassification` specialised for the gptj family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class GPTJForQuestionAnswering(TransformerForQuestionAnswering):
    """GPTJForQuestionAnswering — `TransformerForQuestionAnswering` specialised for the gptj family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'GPTJModel', 'GPTJForCausalLM', 'GPTJForSequenceClassification', 'GPTJForQuestionAnswering']
