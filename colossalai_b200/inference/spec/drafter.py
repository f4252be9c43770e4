"""Greedy drafter for speculative decoding.  Parity: reference `colossalai/inference/spec/drafter.py:13-123`
(`Drafter.speculate(input_ids, n_spec_tokens, past_key_values, glide_input)` with KV trim)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .struct import DrafterOutput, GlideInput

__all__ = ["Drafter"]


class Drafter:
    def __init__(self, model: nn.Module, tokenizer, device: torch.device = None, dtype: torch.dtype = torch.float16):
        self._tokenizer = tokenizer
        self._device = device or torch.device("cpu")
        self._dtype = dtype
        self._drafter_model = model.to(self._device).to(self._dtype).eval()

    def get_model(self) -> nn.Module:
        return self._drafter_model

    @staticmethod
    def trim_kv_cache(past_key_values, invalid_token_num: int):
        if invalid_token_num < 1 or past_key_values is None:
            return past_key_values
        return tuple((k[:, :, :-invalid_token_num], v[:, :, :-invalid_token_num]) for k, v in past_key_values)

    @torch.inference_mode()
    def speculate(self, input_ids: torch.Tensor, n_spec_tokens: int, past_key_values=None,
                  glide_input: Optional[GlideInput] = None) -> DrafterOutput:
        """Greedy draft of `n_spec_tokens` tokens.  The small model is re-run on the growing context (drafters are
        tiny; a dense-KV variant plugs in through `past_key_values` for HF-style models)."""
        assert n_spec_tokens >= 1, "Number of speculated tokens should be >= 1"
        if input_ids.dim() == 1:
            input_ids = input_ids.unsqueeze(0)
        ids = input_ids.to(self._device)
        toks, all_logits = [], []
        use_glide = glide_input is not None and glide_input.glimpse_ready and \
            hasattr(self._drafter_model, "model") and hasattr(self._drafter_model.model.layers[0], "cross_attn")
        for _ in range(n_spec_tokens):
            out = self._drafter_model(input_ids=ids, glide_input=glide_input) if use_glide \
                else self._drafter_model(input_ids=ids)
            logits = out["logits"] if isinstance(out, dict) else out.logits
            logits = logits.view(ids.shape[0], ids.shape[1], -1)[:, -1]
            nxt = logits.argmax(-1)
            toks.append(nxt)
            all_logits.append(logits)
            ids = torch.cat([ids, nxt.view(-1, 1)], dim=1)
        next_tokens = torch.stack(toks, dim=1).squeeze(0)
        return DrafterOutput(speculated_length=n_spec_tokens, logits=torch.stack(all_logits, 1).squeeze(0),
                             next_tokens=next_tokens, past_key_values=None)
