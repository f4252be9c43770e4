"""Qwen3 (per-head q/k RMSNorm, no QKV bias).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/qwen3.py; modeling/qwen3.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "qwen3-8b"
FAMILY_DEFAULTS = {'qk_norm': True}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class Qwen3Model(TransformerBackboneModel):
    """Qwen3Model — `TransformerBackboneModel` specialised for the qwen3 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class Qwen3ForCausalLM(TransformerLMHeadModel):
    """Qwen3ForCausalLM — `TransformerLMHeadModel` specialised for the qwen3 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class Qwen3ForSequenceClassification(TransformerForSequenceClassification):
    """Qwen3ForSequenceClassification — `TransformerForSequenceClassification` specialised for the qwen3 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'Qwen3Model', 'Qwen3ForCausalLM', 'Qwen3ForSequenceClassification']
