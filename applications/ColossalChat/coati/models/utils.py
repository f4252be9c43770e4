"""Log-prob / masking / reward-shaping helpers.  Parity: reference `coati/models/utils.py:1-160`."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def log_probs_from_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """log p(labels) per position: `logits` [B, S, V], `labels` [B, S] -> [B, S] (fp32 log-softmax)."""
    return torch.gather(F.log_softmax(logits.float(), dim=-1), -1, labels.unsqueeze(-1)).squeeze(-1)


def calc_action_log_probs(logits: torch.Tensor, sequences: torch.Tensor, num_actions: int) -> torch.Tensor:
    """Log-probs of the LAST `num_actions` tokens of `sequences` (the generated part)."""
    lp = log_probs_from_logits(logits[:, :-1], sequences[:, 1:])
    return lp[:, -num_actions:]


def calc_masked_log_probs(logits: torch.Tensor, sequences: torch.Tensor, mask: torch.Tensor,
                          length_normalization: bool = False) -> torch.Tensor:
    """Per-token log-probs of `sequences[:, 1:]` zeroed outside `mask[:, 1:]` (optionally divided by the length)."""
    lp = log_probs_from_logits(logits[:, :-1], sequences[:, 1:]) * mask[:, 1:].to(torch.float32)
    if length_normalization:
        lp = lp / mask[:, 1:].sum(-1, keepdim=True).clamp(min=1)
    return lp


def masked_mean(x: torch.Tensor, mask: Optional[torch.Tensor], dim: int = -1) -> torch.Tensor:
    if mask is None:
        return x.mean(dim)
    m = mask.to(x.dtype)
    return (x * m).sum(dim) / m.sum(dim).clamp(min=1e-8)


def masked_whiten(x: torch.Tensor, mask: torch.Tensor, shift_mean: bool = True) -> torch.Tensor:
    m = mask.to(x.dtype)
    mean = (x * m).sum() / m.sum().clamp(min=1)
    var = ((x - mean) ** 2 * m).sum() / m.sum().clamp(min=1)
    out = (x - mean) * torch.rsqrt(var + 1e-8)
    return out if shift_mean else out + mean


def compute_reward(r: torch.Tensor, kl_coef: float, log_probs: torch.Tensor, ref_log_probs: torch.Tensor,
                   action_mask: torch.Tensor, reward_eps: float = 5.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Token-level reward = -kl_coef * (log pi - log pi_ref) everywhere + the (clipped) sequence reward on the last
    action token.  Returns (reward [B, A], mean approximate KL [B])."""
    kl = (log_probs - ref_log_probs) * action_mask
    reward = -kl_coef * kl
    last = action_mask.long().cumsum(-1).argmax(-1)
    reward[torch.arange(r.shape[0], device=r.device), last] += r.clamp(-reward_eps, reward_eps).to(reward.dtype)
    return reward, masked_mean(kl, action_mask, dim=-1)
