"""BERT encoder (post-LN, learned positions + token types, bidirectional).

All classes share the generic parallel-aware backbone (`models/transformer.py`); this module pins the family's
config defaults and exposes the HF-named entry points.  Parity: reference `colossalai/shardformer/policies/bert.py; modeling/bert.py`.
"""
from __future__ import annotations

from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "bert-base"
FAMILY_DEFAULTS = {}


def default_config(**overrides) -> ModelConfig:
    """The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."""
    return get_config(DEFAULT_PRESET, **overrides)


class BertModel(TransformerBackboneModel):
    """BertModel — `TransformerBackboneModel` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForPreTraining(TransformerForMaskedLM):
    """BertForPreTraining — `TransformerForMaskedLM` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertLMHeadModel(TransformerForMaskedLM):
    """BertLMHeadModel — `TransformerForMaskedLM` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForMaskedLM(TransformerForMaskedLM):
    """BertForMaskedLM — `TransformerForMaskedLM` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForSequenceClassification(TransformerForSequenceClassification):
    """BertForSequenceClassification — `TransformerForSequenceClassification` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForTokenClassification(TransformerForTokenClassification):
    """BertForTokenClassification — `TransformerForTokenClassification` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForNextSentencePrediction(TransformerForSequenceClassification):
    """BertForNextSentencePrediction — `TransformerForSequenceClassification` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForMultipleChoice(TransformerForMultipleChoice):
    """BertForMultipleChoice — `TransformerForMultipleChoice` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BertForQuestionAnswering(TransformerForQuestionAnswering):
    """BertForQuestionAnswering — `TransformerForQuestionAnswering` specialised for the bert family."""

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ['default_config', 'BertModel', 'BertForPreTraining', 'BertLMHeadModel', 'BertForMaskedLM', 'BertForSequenceClassification', 'BertForTokenClassification', 'BertForNextSentencePrediction', 'BertForMultipleChoice', 'BertForQuestionAnswering']
