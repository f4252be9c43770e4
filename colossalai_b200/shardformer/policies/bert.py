"""Sharding policies for the bert family.  Parity: reference `colossalai/shardformer/policies/bert.py`.
The family rides on the generic `TransformerPolicy`; subclasses exist so users can override per-head behaviour
(`custom_policy`) exactly like with the reference's per-class policies."""
from __future__ import annotations

from .transformer import TransformerPolicy


class BertModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertModel`."""


class BertForPreTrainingPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForPreTraining`."""


class BertLMHeadModelPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertLMHeadModel`."""


class BertForMaskedLMPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForMaskedLM`."""


class BertForSequenceClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForSequenceClassification`."""


class BertForTokenClassificationPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForTokenClassification`."""


class BertForNextSentencePredictionPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForNextSentencePrediction`."""


class BertForMultipleChoicePolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForMultipleChoice`."""


class BertForQuestionAnsweringPolicy(TransformerPolicy):
    """Policy for `colossalai_b200.models.bert.BertForQuestionAnswering`."""


__all__ = ['BertModelPolicy', 'BertForPreTrainingPolicy', 'BertLMHeadModelPolicy', 'BertForMaskedLMPolicy', 'BertForSequenceClassificationPolicy', 'BertForTokenClassificationPolicy', 'BertForNextSentencePredictionPolicy', 'BertForMultipleChoicePolicy', 'BertForQuestionAnsweringPolicy']
