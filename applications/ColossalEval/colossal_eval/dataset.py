"""Dataset adapters and few-shot prompt construction (reference `colossal_eval/dataset/{mmlu,cmmlu,agieval,gsm,...}.py`,
`colossal_eval/utils/conversation.py`).

Every adapter turns raw rows of a public benchmark layout into the evaluator's item schema
    {"instruction": str, "choices": [str, ...], "answer": int, "category": str}         multiple choice
    {"instruction": str, "target": str, "category": str}                                 free-form generation
so the harness itself never knows about a benchmark.  No network access: callers hand in rows (dicts) or local files."""
from __future__ import annotations

import csv
import json
import re
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence

__all__ = ["mmlu_items", "load_mmlu_csv", "gsm8k_items", "extract_last_number", "cloze_items", "few_shot_prompt",
           "group_by_category", "load_jsonl"]

_LETTERS = "ABCDEFGHIJ"


def load_jsonl(path) -> List[Dict]:
    return [json.loads(line) for line in Path(path).read_text().splitlines() if line.strip()]


def mmlu_items(rows: Iterable[Dict], subject: str = "") -> List[Dict]:
    """rows: {"question", "choices": [...] | "A".."D", "answer": index | letter}.  The instruction ends with "Answer:"
    and the scored continuations are " A" / " B" / ... (option letters, the MMLU protocol), the option texts being part
    of the prompt."""
    items = []
    for r in rows:
        choices = r["choices"] if "choices" in r else [r[l] for l in _LETTERS if l in r]
        ans = r["answer"]
        ans = _LETTERS.index(ans.strip().upper()) if isinstance(ans, str) else int(ans)
        body = r["question"].strip() + "\n" + "\n".join(f"{_LETTERS[i]}. {c}" for i, c in enumerate(choices))
        items.append({"instruction": body + "\nAnswer:", "choices": [f" {_LETTERS[i]}" for i in range(len(choices))],
                      "answer": ans, "category": r.get("subject", subject)})
    return items


def load_mmlu_csv(path, subject: Optional[str] = None) -> List[Dict]:
    """The original MMLU csv layout: question, A, B, C, D, answer-letter (no header)."""
    path = Path(path)
    rows = []
    with path.open(newline="") as f:
        for rec in csv.reader(f):
            if len(rec) < 6:
                continue
            rows.append({"question": rec[0], "choices": rec[1:5], "answer": rec[5]})
    return mmlu_items(rows, subject if subject is not None else path.stem.replace("_test", "").replace("_dev", ""))


def extract_last_number(text: str) -> Optional[str]:
    """GSM8K answer protocol: the last number of the completion (commas and a trailing period stripped)."""
    nums = re.findall(r"-?\d[\d,]*\.?\d*", text)
    if not nums:
        return None
    n = nums[-1].replace(",", "").rstrip(".")
    return n[:-2] if n.endswith(".0") else n


def gsm8k_items(rows: Iterable[Dict]) -> List[Dict]:
    """rows: {"question", "answer": "... #### 42"}; target = the number after '####'."""
    items = []
    for r in rows:
        tgt = r["answer"].split("####")[-1].strip().replace(",", "")
        items.append({"instruction": "Question: " + r["question"].strip() + "\nAnswer:", "target": tgt,
                      "category": "gsm8k", "postprocess": "last_number", "solution": r["answer"].split("####")[0].strip()})
    return items


def cloze_items(rows: Iterable[Dict], category: str = "cloze") -> List[Dict]:
    """rows: {"context", "endings": [...], "label"} (HellaSwag / PIQA style): the endings themselves are scored."""
    return [{"instruction": r["context"].strip(), "choices": [" " + e.strip() for e in r["endings"]],
             "answer": int(r["label"]), "category": r.get("category", category)} for r in rows]


def few_shot_prompt(item: Dict, shots: Sequence[Dict], header: str = "") -> Dict:
    """Prefix `item` with solved examples of the same schema (k-shot); returns a new item."""
    parts = [header.strip()] if header.strip() else []
    for s in shots:
        if "choices" in s:
            parts.append(s["instruction"] + s["choices"][s["answer"]])
        else:
            parts.append(s["instruction"] + " " + (s.get("solution", "") + " " if s.get("solution") else "") + s["target"])
    parts.append(item["instruction"])
    out = dict(item)
    out["instruction"] = "\n\n".join(parts)
    return out


def group_by_category(items: Sequence[Dict]) -> Dict[str, List[Dict]]:
    out: Dict[str, List[Dict]] = {}
    for it in items:
        out.setdefault(it.get("category", ""), []).append(it)
    return out
