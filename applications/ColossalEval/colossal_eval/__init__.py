"""Evaluation harness (reference `applications/ColossalEval/colossal_eval`): dataset adapters produce
`{"instruction", "choices" | "target"}` items; the evaluator scores multiple-choice questions by option
log-likelihood and free-form answers by generation + metric (exact match / F1 / ROUGE-L / perplexity)."""
from .dataset import (cloze_items, extract_last_number, few_shot_prompt, group_by_category, gsm8k_items, load_jsonl,
                      load_mmlu_csv, mmlu_items)
from .evaluate import (Evaluator, bleu, exact_match, f1_score, multiple_choice_accuracy, perplexity, rouge_l,
                       score_choices_by_loglikelihood)

__all__ = ["Evaluator", "exact_match", "f1_score", "rouge_l", "bleu", "perplexity", "multiple_choice_accuracy",
           "score_choices_by_loglikelihood", "mmlu_items", "load_mmlu_csv", "gsm8k_items", "cloze_items",
           "few_shot_prompt", "group_by_category", "extract_last_number", "load_jsonl"]
