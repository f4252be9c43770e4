"""Backbone wrappers: scalar-head models for the critic and the reward model.
Parity: reference `coati/models/{base.py:1-70, critic.py:1-40, reward_model.py:1-50}`."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from colossalai_b200.models import build_model


def get_logits(model: nn.Module, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Causal-LM logits as `[B, S, V]` (our models are token-major internally)."""
    out = model(input_ids=input_ids, attention_mask=attention_mask)
    logits = out["logits"] if isinstance(out, dict) else out.logits
    B, S = input_ids.shape
    V = getattr(getattr(model, "cfg", None), "vocab_size", logits.shape[-1])
    return logits.reshape(B, S, -1)[..., :V]


def disable_dropout(model: nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    cfg = getattr(model, "cfg", None)
    if cfg is not None:
        for f in ("attn_dropout", "hidden_dropout"):
            if hasattr(cfg, f):
                setattr(cfg, f, 0.0)


class BaseModel(nn.Module):
    """Decoder backbone (no LM head) + access to the last hidden states `[B, S, H]`."""

    def __init__(self, pretrained=None, config=None, **kw) -> None:
        super().__init__()
        lm = pretrained if isinstance(pretrained, nn.Module) else build_model(pretrained or config, **kw)
        self.cfg = self.config = lm.cfg
        self.model = lm.model if hasattr(lm, "model") else lm
        self.hidden_size = self.cfg.hidden_size

    def hidden_states(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        h = self.model(input_ids=input_ids, attention_mask=attention_mask)
        h = h["last_hidden_state"] if isinstance(h, dict) else h
        return h.reshape(input_ids.shape[0], input_ids.shape[1], -1)


class RewardModel(BaseModel):
    """Scalar reward read at the LAST non-pad token of every sequence."""

    def __init__(self, pretrained=None, config=None, **kw) -> None:
        super().__init__(pretrained, config, **kw)
        self.value_head = nn.Linear(self.hidden_size, 1)
        nn.init.normal_(self.value_head.weight, std=1.0 / (self.hidden_size + 1))
        nn.init.zeros_(self.value_head.bias)

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        h = self.hidden_states(input_ids, attention_mask)
        if attention_mask is None:
            last = torch.full((input_ids.shape[0],), input_ids.shape[1] - 1, device=input_ids.device)
        else:
            pos = torch.arange(input_ids.shape[1], device=input_ids.device)[None]
            last = (pos * attention_mask.long()).max(dim=1).values
        h_last = h[torch.arange(h.shape[0], device=h.device), last]
        return self.value_head(h_last.to(self.value_head.weight.dtype)).squeeze(-1)


class Critic(BaseModel):
    """Per-token value estimates `[B, S]`."""

    def __init__(self, pretrained=None, config=None, **kw) -> None:
        super().__init__(pretrained, config, **kw)
        self.value_head = nn.Linear(self.hidden_size, 1)
        nn.init.normal_(self.value_head.weight, std=1.0 / (self.hidden_size + 1))
        nn.init.zeros_(self.value_head.bias)

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        h = self.hidden_states(input_ids, attention_mask)
        return self.value_head(h.to(self.value_head.weight.dtype)).squeeze(-1)
