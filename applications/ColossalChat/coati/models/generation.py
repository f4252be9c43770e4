"""Sampling loop used by the experience makers (temperature / top-k / top-p, per-sequence EOS stop, left-padded
prompts).  For throughput-critical rollouts use `coati.distributed.producer` (paged-KV engine); this path re-runs the
model on the growing sequence and works with any wrapped / sharded policy.
Parity: reference `coati/models/generation.py:1-160`."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .base import get_logits


def _filter(logits: torch.Tensor, top_k: Optional[int], top_p: Optional[float]) -> torch.Tensor:
    if top_k:
        kth = logits.topk(min(top_k, logits.shape[-1]), dim=-1).values[..., -1:]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and top_p < 1.0:
        sl, si = logits.sort(descending=True, dim=-1)
        cum = sl.softmax(-1).cumsum(-1)
        drop = cum - sl.softmax(-1) > top_p
        sl = sl.masked_fill(drop, float("-inf"))
        logits = torch.full_like(logits, float("-inf")).scatter(-1, si, sl)
    return logits


def _apply_repetition_penalty(logits: torch.Tensor, seq: torch.Tensor, mask: torch.Tensor, penalty: float) -> torch.Tensor:
    """CTRL-style penalty: logits of tokens already present in the (un-padded) sequence are divided (if positive) or
    multiplied (if negative) by `penalty`."""
    if penalty == 1.0:
        return logits
    seen = torch.zeros_like(logits, dtype=torch.bool).scatter_(1, seq, mask.bool())
    return torch.where(seen, torch.where(logits > 0, logits / penalty, logits * penalty), logits)


def _hit_stop(seq: torch.Tensor, prompt_len: int, stop_sequences) -> torch.Tensor:
    """Which rows end (inside the generated part) with one of the multi-token stop sequences."""
    hit = torch.zeros(seq.shape[0], dtype=torch.bool, device=seq.device)
    for stop in stop_sequences:
        n = len(stop)
        if n == 0 or seq.shape[1] - prompt_len < n:
            continue
        tail = seq[:, -n:]
        hit |= (tail == torch.tensor(stop, device=seq.device)).all(-1)
    return hit


@torch.no_grad()
def generate(model: nn.Module, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
             max_new_tokens: int = 32, do_sample: bool = True, temperature: float = 1.0, top_k: Optional[int] = None,
             top_p: Optional[float] = None, eos_token_id: Optional[int] = None, pad_token_id: int = 0,
             generator: Optional[torch.Generator] = None, repetition_penalty: float = 1.0, min_new_tokens: int = 0,
             stop_sequences=None, return_action_mask: bool = False):
    """Returns `[B, S_prompt + max_new_tokens]` (right-padded with `pad_token_id` after EOS / a stop sequence).
    `min_new_tokens` suppresses EOS until that many tokens exist; `stop_sequences`: lists of token ids that end a
    row like EOS does (the stop sequence itself is kept); `return_action_mask`: also return the bool mask of generated,
    non-padding positions `[B, max_new_tokens]` (what the RL trainers weight their losses with)."""
    seq = input_ids
    prompt_len = input_ids.shape[1]
    mask = attention_mask if attention_mask is not None else torch.ones_like(input_ids)
    done = torch.zeros(seq.shape[0], dtype=torch.bool, device=seq.device)
    was_training = model.training
    model.eval()
    for step in range(max_new_tokens):
        logits = get_logits(model, seq, mask)[:, -1].float()
        logits = _apply_repetition_penalty(logits, seq, mask, repetition_penalty)
        if eos_token_id is not None and step < min_new_tokens:
            logits[:, eos_token_id] = float("-inf")
        if do_sample:
            probs = _filter(logits / max(temperature, 1e-5), top_k, top_p).softmax(-1)
            nxt = torch.multinomial(probs, 1, generator=generator).squeeze(-1)
        else:
            nxt = logits.argmax(-1)
        nxt = torch.where(done, torch.full_like(nxt, pad_token_id), nxt)
        seq = torch.cat([seq, nxt[:, None]], dim=1)
        mask = torch.cat([mask, (~done)[:, None].to(mask.dtype)], dim=1)
        if eos_token_id is not None:
            done = done | (nxt == eos_token_id)
        if stop_sequences:
            done = done | _hit_stop(seq, prompt_len, stop_sequences)
    model.train(was_training)
    if return_action_mask:
        return seq, mask[:, prompt_len:].bool()
    return seq
