"""Token search.  Parity: reference `colossalai/inference/sampler.py` (greedy / multinomial / beam placeholder)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .logit_processors import get_logits_processor

__all__ = ["greedy_sample", "multinomial_sample", "beam_search_sample", "search_tokens"]


def greedy_sample(logprobs: torch.Tensor) -> torch.Tensor:
    return torch.argmax(logprobs, dim=-1)


def multinomial_sample(probs: torch.Tensor) -> torch.Tensor:
    return torch.multinomial(probs, num_samples=1, replacement=True).squeeze(1)


def beam_search_sample(beam_width: int, logprobs: torch.Tensor, is_prompt: bool = False,
                       cumulative_logprobs: Optional[torch.Tensor] = None) -> List[Tuple[List[int], List[int]]]:
    """Per sequence group: (parent beam ids, next token ids).  logprobs: [n_beams_total, vocab]."""
    results = []
    if is_prompt:
        for row in logprobs:
            _, next_ids = torch.topk(row, beam_width)
            results.append(([0] * beam_width, next_ids.tolist()))
        return results
    n_groups = logprobs.shape[0] // beam_width
    for g in range(n_groups):
        lp = logprobs[g * beam_width:(g + 1) * beam_width]
        if cumulative_logprobs is not None:
            lp = lp + cumulative_logprobs[g * beam_width:(g + 1) * beam_width].unsqueeze(1)
        flat = lp.flatten()
        _, top = torch.topk(flat, beam_width)
        vocab = lp.shape[-1]
        results.append(((top // vocab).tolist(), (top % vocab).tolist()))
    return results


def search_tokens(generation_config, logits: torch.Tensor, is_prompt: bool = False,
                  batch_token_ids: Optional[List[List[int]]] = None, sequence_lengths=None, max_lengths=None,
                  eos_token_id=None) -> torch.Tensor:
    """Apply logit processors then sample one token per row of `logits` [bsz, vocab]."""
    cfg = generation_config.to_dict() if hasattr(generation_config, "to_dict") else dict(generation_config)
    logits = logits.float()
    for t in ("no_repeat_ngram_size", "repetition_penalty"):
        if cfg.get(t) is not None and batch_token_ids is not None:
            logits = get_logits_processor(t, logits, cfg[t], batch_token_ids)
    if cfg.get("forced_eos_token_id") is not None and sequence_lengths is not None and max_lengths is not None:
        logits = get_logits_processor("forced_eos_token_id", logits, sequence_lengths, max_lengths,
                                      cfg["forced_eos_token_id"])
    if cfg.get("do_sample"):
        for t in ("temperature", "top_k", "top_p"):
            if cfg.get(t) is not None:
                logits = get_logits_processor(t, logits, cfg[t])
        return multinomial_sample(torch.softmax(logits, dim=-1))
    if cfg.get("num_beams", 1) > 1:
        raise NotImplementedError("beam search is exposed through beam_search_sample(); the engine decodes greedily")
    return greedy_sample(torch.log_softmax(logits, dim=-1))
