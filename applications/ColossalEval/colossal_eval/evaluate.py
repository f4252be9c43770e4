"""Parity: reference `colossal_eval/{models/huggingface.py (batched loss / generation inference),
evaluate/dataset_evaluator/{dataset_evaluator.py, metrics.py}}`."""
from __future__ import annotations

import math
import re
import string
from collections import Counter
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def _norm(s: str) -> str:
    s = s.lower()
    s = "".join(ch for ch in s if ch not in set(string.punctuation))
    s = re.sub(r"\b(a|an|the)\b", " ", s)
    return " ".join(s.split())


def exact_match(pred: str, ref: str) -> float:
    return float(_norm(pred) == _norm(ref))


def f1_score(pred: str, ref: str) -> float:
    p, r = _norm(pred).split(), _norm(ref).split()
    common = sum((Counter(p) & Counter(r)).values())
    if common == 0:
        return 0.0
    prec, rec = common / len(p), common / len(r)
    return 2 * prec * rec / (prec + rec)


def rouge_l(pred: str, ref: str) -> float:
    a, b = _norm(pred).split(), _norm(ref).split()
    if not a or not b:
        return 0.0
    dp = [[0] * (len(b) + 1) for _ in range(len(a) + 1)]
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            dp[i + 1][j + 1] = dp[i][j] + 1 if x == y else max(dp[i][j + 1], dp[i + 1][j])
    l = dp[-1][-1]
    if l == 0:
        return 0.0
    p, r = l / len(a), l / len(b)
    return 2 * p * r / (p + r)


@torch.no_grad()
def _token_logprobs(model, ids: torch.Tensor) -> torch.Tensor:
    logits = model(input_ids=ids)["logits"].reshape(ids.shape[0], ids.shape[1], -1)[..., : model.cfg.vocab_size]
    return torch.gather(F.log_softmax(logits[:, :-1].float(), -1), -1, ids[:, 1:, None]).squeeze(-1)


@torch.no_grad()
def score_choices_by_loglikelihood(model, tokenizer: Callable, prompt: str, choices: Sequence[str],
                                   length_normalize: bool = True) -> List[float]:
    dev = next(model.parameters()).device
    p = list(tokenizer(prompt))
    scores = []
    for c in choices:
        ct = list(tokenizer(c))
        lp = _token_logprobs(model, torch.tensor([p + ct], device=dev))[0, len(p) - 1:]
        scores.append(float(lp.sum() / (len(ct) if length_normalize else 1)))
    return scores


@torch.no_grad()
def perplexity(model, tokenizer: Callable, texts: Sequence[str]) -> float:
    dev = next(model.parameters()).device
    nll, n = 0.0, 0
    for t in texts:
        ids = torch.tensor([list(tokenizer(t))], device=dev)
        lp = _token_logprobs(model, ids)
        nll -= float(lp.sum())
        n += lp.numel()
    return math.exp(nll / max(n, 1))


def multiple_choice_accuracy(model, tokenizer: Callable, items: Sequence[Dict]) -> float:
    hit = 0
    for it in items:
        s = score_choices_by_loglikelihood(model, tokenizer, it["instruction"], it["choices"])
        hit += int(max(range(len(s)), key=s.__getitem__) == it["answer"])
    return hit / max(1, len(items))


class Evaluator:
    """`Evaluator(model, tokenizer).evaluate({"mmlu-like": items, "qa": items}, metrics={"qa": ["exact_match", "f1"]})`."""

    METRICS = {"exact_match": exact_match, "f1": f1_score, "rouge_l": rouge_l}

    def __init__(self, model, tokenizer: Callable, decode: Optional[Callable] = None, max_new_tokens: int = 32,
                 eos_token_id: int = 2) -> None:
        self.model, self.tokenizer, self.decode = model.eval(), tokenizer, decode
        self.max_new_tokens, self.eos_token_id = max_new_tokens, eos_token_id

    @torch.no_grad()
    def generate(self, prompt: str) -> str:
        dev = next(self.model.parameters()).device
        ids = torch.tensor([list(self.tokenizer(prompt))], device=dev)
        out: List[int] = []
        for _ in range(self.max_new_tokens):
            lg = self.model(input_ids=ids)["logits"].reshape(1, ids.shape[1], -1)[0, -1, : self.model.cfg.vocab_size]
            nxt = int(lg.argmax())
            if nxt == self.eos_token_id:
                break
            out.append(nxt)
            ids = torch.cat([ids, torch.tensor([[nxt]], device=dev)], 1)
        return self.decode(out) if self.decode else " ".join(map(str, out))

    def evaluate(self, datasets: Dict[str, Sequence[Dict]], metrics: Optional[Dict[str, List[str]]] = None) -> Dict[str, Dict[str, float]]:
        res: Dict[str, Dict[str, float]] = {}
        for name, items in datasets.items():
            if items and "choices" in items[0]:
                res[name] = {"accuracy": multiple_choice_accuracy(self.model, self.tokenizer, items)}
                continue
            names = (metrics or {}).get(name, ["exact_match"])
            agg = {m: 0.0 for m in names}
            for it in items:
                pred = self.generate(it["instruction"])
                for m in names:
                    agg[m] += self.METRICS[m](pred, it["target"])
            res[name] = {m: v / max(1, len(items)) for m, v in agg.items()}
        return res
