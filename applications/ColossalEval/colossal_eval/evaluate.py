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
def bleu(pred: str, ref: str, max_n: int = 4) -> float:
    """Sentence BLEU (uniform n-gram weights up to `max_n`, brevity penalty, +1 smoothing above unigrams)."""
    import math
    from collections import Counter

    p, r = _norm(pred).split(), _norm(ref).split()
    if not p or not r:
        return 0.0
    logs = 0.0
    for n in range(1, max_n + 1):
        pn = Counter(tuple(p[i:i + n]) for i in range(len(p) - n + 1))
        rn = Counter(tuple(r[i:i + n]) for i in range(len(r) - n + 1))
        hit = sum(min(c, rn[g]) for g, c in pn.items())
        tot = max(1, sum(pn.values()))
        if n == 1:
            if hit == 0:
                return 0.0
            logs += math.log(hit / tot)
        else:
            logs += math.log((hit + 1) / (tot + 1))
    bp = 1.0 if len(p) > len(r) else math.exp(1 - len(r) / len(p))
    return bp * math.exp(logs / max_n)


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

    METRICS = {"exact_match": exact_match, "f1": f1_score, "rouge_l": rouge_l, "bleu": bleu}

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
                if it.get("postprocess") == "last_number":                 # GSM8K protocol
                    from .dataset import extract_last_number

                    pred = extract_last_number(pred) or ""
                for m in names:
                    agg[m] += self.METRICS[m](pred, it["target"])
            res[name] = {m: v / max(1, len(items)) for m, v in agg.items()}
        return res

    def evaluate_by_category(self, items: Sequence[Dict], metrics: Optional[List[str]] = None,
                             shots: Optional[Sequence[Dict]] = None, header: str = "") -> Dict[str, Dict[str, float]]:
        """One benchmark with a `category` per item (MMLU subjects, ...): per-category scores, the macro average over
        categories and the micro average over items; `shots` are prepended to every item (k-shot prompting)."""
        from .dataset import few_shot_prompt, group_by_category

        if shots:
            items = [few_shot_prompt(it, shots, header) for it in items]
        groups = group_by_category(items)
        per = self.evaluate(groups, {k: (metrics or ["exact_match"]) for k in groups})
        keys = sorted({m for v in per.values() for m in v})
        total = sum(len(g) for g in groups.values())
        per["macro_avg"] = {m: sum(per[c][m] for c in groups) / max(1, len(groups)) for m in keys}
        per["micro_avg"] = {m: sum(per[c][m] * len(groups[c]) for c in groups) / max(1, total) for m in keys}
        return per

    def save(self, results: Dict, path) -> None:
        import json
        from pathlib import Path

        Path(path).parent.mkdir(parents=True, exist_ok=True)
        Path(path).write_text(json.dumps(results, indent=2, sort_keys=True))
