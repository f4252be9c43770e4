"""Evaluation harness (reference `applications/ColossalEval/colossal_eval`): dataset adapters produce
`{"instruction", "choices" | "target"}` items; the evaluator scores multiple-choice questions by option
log-likelihood and free-form answers by generation + metric (exact match / F1 / ROUGE-L / perplexity)."""
from .evaluate import (Evaluator, exact_match, f1_score, multiple_choice_accuracy, perplexity, rouge_l,
                       score_choices_by_loglikelihood)

__all__ = ["Evaluator", "exact_match", "f1_score", "rouge_l", "perplexity", "multiple_choice_accuracy",
           "score_choices_by_loglikelihood"]
