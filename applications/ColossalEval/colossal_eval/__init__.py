"""Evaluation harness (reference `applications/ColossalEval/colossal_eval`): dataset adapters produce
`{"instruction", "choices" | "target"}` items; the evaluator scores multiple-choice questions by option
log-likelihood and free-form answers by generation + metric (exact match / F1 / ROUGE-L / perplexity).  `EvalModel` +
`run_inference` / `run_evaluation` are the batched, multi-rank, two-phase pipeline (answers saved to json, scored
offline); `judge` is the LLM-as-judge ("GPT evaluation") path with a pluggable judge callable."""
from .dataset import (cloze_items, extract_last_number, few_shot_prompt, group_by_category, gsm8k_items, load_jsonl,
                      load_mmlu_csv, mmlu_items)
from .judge import CRITERIA, judge_battle, judge_scores, local_judge, parse_battle, parse_score
from .models import EvalModel
from .pipeline import format_table, run_evaluation, run_inference
from .evaluate import (Evaluator, bleu, exact_match, f1_score, multiple_choice_accuracy, perplexity, rouge_l,
                       score_choices_by_loglikelihood)

__all__ = ["EvalModel", "run_inference", "run_evaluation", "format_table", "judge_scores", "judge_battle", "local_judge",
           "parse_score", "parse_battle", "CRITERIA", "Evaluator", "exact_match", "f1_score", "rouge_l", "bleu", "perplexity", "multiple_choice_accuracy",
           "score_choices_by_loglikelihood", "mmlu_items", "load_mmlu_csv", "gsm8k_items", "cloze_items",
           "few_shot_prompt", "group_by_category", "extract_last_number", "load_jsonl"]
