"""DeepSeek-V3 MoE routing (sigmoid scores, group-limited top-k, routed scaling) on the generic block.

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/deepseek_v3.py; modeling/deepseek_v3.py:26`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "deepseek-tiny"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class DeepseekV3Model(TransformerBackboneModel):
    """DeepseekV3Model — `TransformerBackboneModel` specialised for the deepseek_v3 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class DeepseekV3ForCausalLM(TransformerLMHeadModel):
    """DeepseekV3ForCausalLM — `TransformerLMHeadModel` specialised for the deepseek_v3 family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'DeepseekV3Model', 'DeepseekV3ForCausalLM']
