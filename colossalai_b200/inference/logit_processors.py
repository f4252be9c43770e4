"""Logit processors.  Parity: reference `colossalai/inference/logit_processors.py` (no-repeat-ngram, repetition
penalty, temperature, top-k, top-p, forced EOS) with the same registry API."""
from __future__ import annotations

from typing import Dict, List, Union

import torch
import torch.nn.functional as F

__all__ = ["register_logits_processor", "get_logits_processor"]

_LOGITS_PROCESSOR_MAP: Dict[str, callable] = {}


def register_logits_processor(name: str):
    def deco(fn):
        _LOGITS_PROCESSOR_MAP[name] = fn
        return fn

    return deco


@register_logits_processor("no_repeat_ngram_size")
def apply_no_repeat_ngram_size(logits: torch.Tensor, ngram_size: int, batch_token_ids: List[List[int]]) -> torch.Tensor:
    if not isinstance(ngram_size, int) or ngram_size < 0:
        raise ValueError(f"'ngram_size={ngram_size}' should be a non-negative integer.")
    if ngram_size != 0:
        for b, ids in enumerate(batch_token_ids):
            if len(ids) + 1 < ngram_size:
                continue
            grams: Dict[tuple, List[int]] = {}
            for i in range(len(ids) - ngram_size + 1):
                g = tuple(ids[i:i + ngram_size])
                grams.setdefault(g[:-1], []).append(g[-1])
            prefix = tuple(ids[len(ids) - ngram_size + 1:]) if ngram_size > 1 else ()
            for tok in grams.get(prefix, []):
                logits[b, tok] = -float("inf")
    return logits


@register_logits_processor("repetition_penalty")
def apply_repetition_penalty(logits: torch.Tensor, penalty: float, batch_token_ids: List[List[int]]) -> torch.Tensor:
    if not isinstance(penalty, float) or not (penalty > 0):
        raise ValueError(f"'penalty={penalty}' has to be a strictly positive float and greater than 0.")
    if penalty != 1.0:
        for b, ids in enumerate(batch_token_ids):
            if not ids:
                continue
            idx = torch.tensor(sorted(set(ids)), device=logits.device, dtype=torch.long)
            row = logits[b, idx]
            logits[b, idx] = torch.where(row > 0, row / penalty, row * penalty)
    return logits


@register_logits_processor("temperature")
def apply_temperature(logits: torch.Tensor, temperature: float) -> torch.Tensor:
    if not isinstance(temperature, float) or not (0.0 < temperature <= 1.0):
        raise ValueError(f"'temperature={temperature}' should be a strictly positive float, less than or equal to 1.0")
    return logits if temperature == 1.0 else logits / temperature


@register_logits_processor("top_k")
def apply_top_k(logits: torch.Tensor, top_k: int) -> torch.Tensor:
    if not isinstance(top_k, int) or top_k <= 0:
        raise ValueError(f"`top_k` should be a strictly positive integer, but got {top_k}.")
    top_k = min(top_k, logits.size(-1))
    kth = torch.topk(logits, top_k)[0][..., -1, None]
    return logits.masked_fill(logits < kth, -float("inf"))


@register_logits_processor("top_p")
def apply_top_p(logits: torch.Tensor, top_p: float) -> torch.Tensor:
    if top_p < 0 or top_p > 1.0:
        raise ValueError(f"`top_p` should be a float > 0 and < 1, but got {top_p}.")
    if top_p >= 1.0:
        return logits
    sorted_logits, sorted_idx = torch.sort(logits, descending=True)
    cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
    remove = cum > top_p
    remove[..., 1:] = remove[..., :-1].clone()
    remove[..., 0] = False
    mask = remove.scatter(dim=-1, index=sorted_idx, src=remove)
    return logits.masked_fill(mask, -float("inf"))


@register_logits_processor("forced_eos_token_id")
def apply_forced_eos_token_id(logits: torch.Tensor, sequence_lengths: Union[torch.Tensor, List[int]],
                              max_lengths: Union[torch.Tensor, List[int]],
                              eos_token_id: Union[int, List[int]]) -> torch.Tensor:
    if isinstance(sequence_lengths, torch.Tensor):
        sequence_lengths = sequence_lengths.tolist()
    if isinstance(max_lengths, torch.Tensor):
        max_lengths = max_lengths.tolist()
    if isinstance(eos_token_id, int):
        eos_token_id = [eos_token_id]
    for b, (sl, ml) in enumerate(zip(sequence_lengths, max_lengths)):
        if sl == ml - 1:
            logits[b, :] = -float("inf")
            logits[b, eos_token_id] = 0
    return logits


def get_logits_processor(processor: str, logits: torch.Tensor, *args, **kwargs) -> torch.Tensor:
    if processor not in _LOGITS_PROCESSOR_MAP:
        return logits
    return _LOGITS_PROCESSOR_MAP[processor](logits, *args, **kwargs)
