"""Verifiable rewards for RLVR (math answers in \\boxed{}, response format).
Parity: reference `coati/distributed/reward/{reward_fn.py, reward_utils.py, verifiable_reward.py}`."""
from __future__ import annotations

import re
from typing import Callable, List, Optional, Sequence

import torch


def extract_boxed(text: str) -> Optional[str]:
    """Content of the LAST `\\boxed{...}` (brace-balanced)."""
    i = text.rfind("\\boxed{")
    if i < 0:
        return None
    depth, j = 0, i + len("\\boxed{") - 1
    for k in range(j, len(text)):
        if text[k] == "{":
            depth += 1
        elif text[k] == "}":
            depth -= 1
            if depth == 0:
                return text[j + 1:k].strip()
    return None


def _normalise(ans: str) -> str:
    ans = ans.strip().replace(" ", "").replace("\\!", "").replace("\\,", "").rstrip(".")
    ans = re.sub(r"\\text\{([^}]*)\}", r"\1", ans)
    ans = re.sub(r"^\\\((.*)\\\)$", r"\1", ans)
    try:
        f = float(ans.replace(",", ""))
        return str(int(f)) if f == int(f) else repr(f)
    except ValueError:
        return ans


def boxed_math_reward(response: str, gt_answer: str, format_score: float = 0.0, correct_score: float = 1.0) -> float:
    """`correct_score` if the boxed answer matches the ground truth, `format_score` if there is a boxed answer at all."""
    got = extract_boxed(response)
    if got is None:
        return 0.0
    return correct_score if _normalise(got) == _normalise(str(gt_answer)) else format_score


def format_reward(response: str, think_tags: Sequence[str] = ("<think>", "</think>"),
                  answer_tags: Sequence[str] = ("<answer>", "</answer>")) -> float:
    """1.0 when the response is `<think>..</think><answer>..</answer>` with each tag exactly once and in order."""
    pos = []
    for tag in (*think_tags, *answer_tags):
        if response.count(tag) != 1:
            return 0.0
        pos.append(response.find(tag))
    return 1.0 if pos == sorted(pos) else 0.0


def make_reward_fn(decode: Callable[[List[int]], str], scorer: Callable[..., float] = boxed_math_reward,
                   eos_token_id: Optional[int] = None) -> Callable:
    """Adapter to the trainer signature `reward_fn(sequences, prompt_len, gt_answer=[...]) -> [B]`."""

    def fn(sequences: torch.Tensor, prompt_len: int, gt_answer: Optional[List[str]] = None, **_) -> torch.Tensor:
        out = []
        for i, row in enumerate(sequences.tolist()):
            resp = row[prompt_len:]
            if eos_token_id is not None and eos_token_id in resp:
                resp = resp[: resp.index(eos_token_id)]
            text = decode(resp)
            out.append(scorer(text, gt_answer[i]) if gt_answer is not None else scorer(text))
        return torch.tensor(out, dtype=torch.float32)

    return fn


def length_penalty(num_tokens: int, max_new_tokens: int, soft_cache: int = 0, penalty: float = 1.0) -> float:
    """DAPO "soft overlong punishment": 0 inside `max_new_tokens - soft_cache`, linearly down to `-penalty` at the
    generation limit (an answer truncated by the limit is penalised whatever its content)."""
    if soft_cache <= 0:
        return -penalty if num_tokens >= max_new_tokens else 0.0
    start = max_new_tokens - soft_cache
    if num_tokens <= start:
        return 0.0
    return -penalty * min(1.0, (num_tokens - start) / soft_cache)


def code_reward(response: str, tests: str, timeout_s: float = 5.0) -> float:
    """1.0 when the LAST fenced python block of the response passes `tests` (python source with asserts) in a fresh
    interpreter with a wall-clock limit; 0.0 otherwise (no block, exception, timeout)."""
    import subprocess
    import sys
    import tempfile

    blocks = re.findall(r"```(?:python)?\n(.*?)```", response, flags=re.S)
    if not blocks:
        return 0.0
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(blocks[-1] + "\n\n" + tests + "\n")
        path = f.name
    try:
        r = subprocess.run([sys.executable, "-I", path], capture_output=True, timeout=timeout_s)
        return 1.0 if r.returncode == 0 else 0.0
    except subprocess.TimeoutExpired:
        return 0.0
    finally:
        import os

        os.unlink(path)


def combine_rewards(*scorers_and_weights) -> Callable[..., float]:
    """`combine_rewards((boxed_math_reward, 1.0), (format_reward, 0.1))` -> scorer(text, gt) = weighted sum; scorers
    that do not take a ground truth are called with the text only."""
    import inspect

    items = []
    for fn, w in scorers_and_weights:
        n = len([p for p in inspect.signature(fn).parameters.values() if p.default is inspect._empty])
        items.append((fn, float(w), n))

    def scorer(text: str, gt: Optional[str] = None) -> float:
        total = 0.0
        for fn, w, n in items:
            total += w * (fn(text, gt) if n >= 2 else fn(text))
        return total

    return scorer
