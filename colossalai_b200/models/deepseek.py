"""DeepSeekMoE (fine-grained routed experts + shared experts, leading dense layers).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/deepseek.py; modeling/deepseek.py:63-230`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "deepseek-moe-16b"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class DeepseekModel(TransformerBackboneModel):
    """DeepseekModel — `TransformerBackboneModel` specialised for the deepseek family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class DeepseekForCausalLM(TransformerLMHeadModel):
    """DeepseekForCausalLM — `TransformerLMHeadModel` specialised for the deepseek family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'DeepseekModel', 'DeepseekForCausalLM']
