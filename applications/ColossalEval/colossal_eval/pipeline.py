"""The two phases of an evaluation run, decoupled through files like the reference's `examples/dataset_evaluation`:

  1. `run_inference(eval_model, datasets, out_dir)`: every rank answers its share of every dataset, the shares are
     merged on rank 0 and written to `<out_dir>/<dataset>_inference.json` (inputs, model outputs, losses) - the
     expensive, GPU-side phase;
  2. `run_evaluation(inference_dir, metrics, out_path)`: scores the saved answers offline (no model): accuracy for
     multiple-choice items, text metrics for generated answers (optionally after a post-processor such as GSM8K's last
     number), perplexity for loss datasets, per category with macro / micro averages, one results json + a table.

Parity: reference `examples/dataset_evaluation/{inference.py, eval_dataset.py}` and
`colossal_eval/evaluate/dataset_evaluator/{dataset_evaluator.py, metrics.py}`.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Sequence

import torch.distributed as dist

from .dataset import extract_last_number, group_by_category
from .evaluate import bleu, exact_match, f1_score, rouge_l

__all__ = ["run_inference", "run_evaluation", "format_table", "POSTPROCESS", "TEXT_METRICS"]

TEXT_METRICS = {"exact_match": exact_match, "f1": f1_score, "rouge_l": rouge_l, "bleu": bleu}
POSTPROCESS = {
    None: lambda s: s,
    "last_number": lambda s: extract_last_number(s) or "",
    "first_line": lambda s: s.strip().splitlines()[0] if s.strip() else "",
    "first_capital_letter": lambda s: next((c for c in s if c in "ABCDEFGH"), ""),
}


def run_inference(eval_model, datasets: Dict[str, Sequence[Dict]], out_dir: str, group=None) -> Dict[str, str]:
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    paths: Dict[str, str] = {}
    for name, items in datasets.items():
        answers = eval_model.inference(items, rank, world)
        if multi:
            box: List[Optional[List[Dict]]] = [None] * world
            dist.all_gather_object(box, answers, group=group)
            answers = sorted((a for share in box for a in share), key=lambda it: it["index"])
        paths[name] = os.path.join(out_dir, f"{name}_inference.json")
        if rank == 0:
            os.makedirs(out_dir, exist_ok=True)
            with open(paths[name], "w") as f:
                json.dump({"dataset": name, "num_items": len(answers), "items": answers}, f, indent=1)
    if multi:
        dist.barrier(group=group)
    return paths


def _score(items: Sequence[Dict], metrics: Sequence[str]) -> Dict[str, float]:
    if not items:
        return {}
    if "choices" in items[0]:
        return {"accuracy": sum(int(it["output"] == it["answer"]) for it in items) / len(items)}
    if "loss" in items[0]:
        tokens = sum(it.get("num_target_tokens", 1) for it in items)
        nll = sum(it["loss"] * it.get("num_target_tokens", 1) for it in items)
        return {"loss": nll / max(tokens, 1), "perplexity": math.exp(min(nll / max(tokens, 1), 50.0))}
    agg = {m: 0.0 for m in metrics}
    for it in items:
        pred = POSTPROCESS[it.get("postprocess")](it["output"])
        for m in metrics:
            agg[m] += TEXT_METRICS[m](pred, it["target"])
    return {m: v / len(items) for m, v in agg.items()}


def run_evaluation(inference_dir: str, metrics: Optional[Dict[str, List[str]]] = None, out_path: Optional[str] = None
                   ) -> Dict[str, Dict[str, Dict[str, float]]]:
    """{dataset: {category or "macro_avg" / "micro_avg": {metric: value}}} from the files `run_inference` wrote."""
    results: Dict[str, Dict[str, Dict[str, float]]] = {}
    for fname in sorted(os.listdir(inference_dir)):
        if not fname.endswith("_inference.json"):
            continue
        with open(os.path.join(inference_dir, fname)) as f:
            blob = json.load(f)
        name, items = blob["dataset"], blob["items"]
        wanted = (metrics or {}).get(name, ["exact_match"])
        groups = group_by_category(items)
        per = {cat: _score(rows, wanted) for cat, rows in groups.items()}
        keys = sorted({m for v in per.values() for m in v})
        total = sum(len(rows) for rows in groups.values())
        per["macro_avg"] = {m: sum(per[c].get(m, 0.0) for c in groups) / max(1, len(groups)) for m in keys}
        per["micro_avg"] = {m: sum(per[c].get(m, 0.0) * len(groups[c]) for c in groups) / max(1, total) for m in keys}
        results[name] = per
    if out_path:
        os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(results, f, indent=2, sort_keys=True)
    return results


def format_table(results: Dict[str, Dict[str, Dict[str, float]]]) -> str:
    lines = [f"{'dataset':18s} {'category':22s} {'metric':12s} {'value':>8s}"]
    for ds, per in results.items():
        for cat in sorted(per, key=lambda c: (c in ("macro_avg", "micro_avg"), c)):
            for m, v in sorted(per[cat].items()):
                lines.append(f"{ds:18s} {cat:22s} {m:12s} {v:8.4f}")
    return "\n".join(lines)
