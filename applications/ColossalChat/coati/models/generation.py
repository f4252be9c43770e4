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


@torch.no_grad()
def generate(model: nn.Module, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
             max_new_tokens: int = 32, do_sample: bool = True, temperature: float = 1.0, top_k: Optional[int] = None,
             top_p: Optional[float] = None, eos_token_id: Optional[int] = None, pad_token_id: int = 0,
             generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Returns `[B, S_prompt + max_new_tokens]` (right-padded with `pad_token_id` after EOS)."""
    seq = input_ids
    mask = attention_mask if attention_mask is not None else torch.ones_like(input_ids)
    done = torch.zeros(seq.shape[0], dtype=torch.bool, device=seq.device)
    was_training = model.training
    model.eval()
    for _ in range(max_new_tokens):
        logits = get_logits(model, seq, mask)[:, -1].float()
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
    model.train(was_training)
    return seq
