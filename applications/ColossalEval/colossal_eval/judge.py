"""LLM-as-judge evaluation of open-ended answers (the reference's "GPT evaluation"): a judge model scores an answer on
named criteria from 1 to 5, or compares two assistants' answers pairwise ("battle"), from a fixed prompt template.

The judge is any callable `str -> str` - an OpenAI-compatible client, the `/chat` endpoint of
`colossalai_b200.inference.server`, or a local model wrapped by `local_judge(EvalModel)`; nothing here needs the
network.  Parsing is defensive: a reply without a parsable score counts as a failure and is reported, not guessed.

Parity: reference `colossal_eval/evaluate/{gpt_evaluate.py:1-850, evaluator.py, utils.py}` and
`configs/gpt_evaluation/prompt/*` (criteria prompts, battle prompt, score extraction, per-category aggregation).
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Optional, Sequence

__all__ = ["CRITERIA", "score_prompt", "battle_prompt", "parse_score", "parse_battle", "judge_scores", "judge_battle",
           "local_judge"]

CRITERIA: Dict[str, str] = {
    "correctness": "Is the answer factually and logically correct?",
    "relevance": "Does the answer address the question that was asked, without digressing?",
    "language organization": "Is the answer fluent, coherent and well structured?",
    "creativity": "Does the answer show original, imaginative ideas where the task invites them?",
    "conciseness": "Is the answer free of redundant content?",
}


def score_prompt(question: str, answer: str, criterion: str, reference: Optional[str] = None) -> str:
    ref = f"\n[Reference answer]\n{reference}\n" if reference else ""
    return (f"You are a strict, impartial evaluator.  Rate the assistant's answer on the criterion "
            f"\"{criterion}\" ({CRITERIA.get(criterion, criterion)}) with an integer from 1 (very poor) to 5 (excellent).\n"
            f"[Question]\n{question}\n{ref}[Assistant's answer]\n{answer}\n\n"
            f"Think step by step in one short paragraph, then finish with a line of the form `Score: <1-5>`.")


def battle_prompt(question: str, answer_a: str, answer_b: str) -> str:
    return ("You are a strict, impartial evaluator comparing two AI assistants.\n"
            f"[Question]\n{question}\n[Assistant A]\n{answer_a}\n[Assistant B]\n{answer_b}\n\n"
            "Explain briefly which answer is better, then finish with exactly one line: `Winner: A`, `Winner: B` or "
            "`Winner: tie`.")


def parse_score(reply: str) -> Optional[int]:
    m = re.findall(r"score\s*[:=]\s*([1-5])(?:\s*/\s*5)?\b", reply, flags=re.I)
    return int(m[-1]) if m else None


def parse_battle(reply: str) -> Optional[str]:
    m = re.findall(r"winner\s*[:=]\s*(a|b|tie)\b", reply, flags=re.I)
    return m[-1].lower() if m else None


def judge_scores(judge: Callable[[str], str], items: Sequence[Dict], criteria: Sequence[str] = ("correctness", "relevance"),
                 retries: int = 1) -> Dict:
    """`items`: {"instruction", "output", optional "target", optional "category"}.  Returns per-item scores, the mean per
    criterion overall and per category, and the number of replies that could not be parsed."""
    rows: List[Dict] = []
    failed = 0
    for it in items:
        scores: Dict[str, Optional[int]] = {}
        for c in criteria:
            s = None
            for _ in range(retries + 1):
                s = parse_score(judge(score_prompt(it["instruction"], it["output"], c, it.get("target"))))
                if s is not None:
                    break
            failed += s is None
            scores[c] = s
        rows.append({"instruction": it["instruction"], "category": it.get("category", ""), "scores": scores})

    def mean(sel, c):
        v = [r["scores"][c] for r in sel if r["scores"][c] is not None]
        return sum(v) / len(v) if v else float("nan")

    cats = sorted({r["category"] for r in rows})
    return {"items": rows, "failed": failed,
            "overall": {c: mean(rows, c) for c in criteria},
            "by_category": {k: {c: mean([r for r in rows if r["category"] == k], c) for c in criteria} for k in cats}}


def judge_battle(judge: Callable[[str], str], questions: Sequence[str], answers_a: Sequence[str],
                 answers_b: Sequence[str], swap: bool = True) -> Dict[str, float]:
    """Pairwise comparison; with `swap` every pair is judged in both orders (position bias) and a pair only counts
    as a win when both orders agree, otherwise as a tie."""
    tally = {"a": 0, "b": 0, "tie": 0, "failed": 0}
    for q, a, b in zip(questions, answers_a, answers_b):
        first = parse_battle(judge(battle_prompt(q, a, b)))
        verdict = first
        if swap and first is not None:
            second = parse_battle(judge(battle_prompt(q, b, a)))
            second = {"a": "b", "b": "a", "tie": "tie"}.get(second) if second is not None else None
            verdict = first if second == first else ("tie" if second is not None else None)
        tally[verdict if verdict is not None else "failed"] += 1
    n = max(1, len(questions) - tally["failed"])
    return {"win_rate_a": tally["a"] / n, "win_rate_b": tally["b"] / n, "tie_rate": tally["tie"] / n,
            "failed": tally["failed"], "n": len(questions)}


def local_judge(eval_model, max_new_tokens: int = 128) -> Callable[[str], str]:
    """A judge backed by a local `EvalModel` (greedy decoding of the evaluation prompt)."""
    return lambda prompt: eval_model.generate([prompt], max_new_tokens=max_new_tokens)[0]
