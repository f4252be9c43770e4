"""Rollout storage + the PPO experience maker (generate -> log-probs of actor and reference -> reward model ->
KL-shaped token rewards -> GAE).
Parity: reference `coati/experience_buffer/{base,naive,utils}.py` and `coati/experience_maker/{base,naive}.py`."""
from __future__ import annotations

import random
from dataclasses import dataclass, fields
from typing import Callable, List, Optional

import torch
import torch.nn as nn

from .models import calc_action_log_probs, compute_reward, generate, get_logits


@dataclass
class Experience:
    sequences: torch.Tensor            # [B, S]
    action_log_probs: torch.Tensor     # [B, A]
    values: torch.Tensor               # [B, A]
    reward: torch.Tensor               # [B]      sequence-level reward (before KL shaping)
    kl: torch.Tensor                   # [B]
    advantages: torch.Tensor           # [B, A]
    attention_mask: torch.Tensor       # [B, S]
    action_mask: torch.Tensor          # [B, A]

    def to_device(self, device) -> "Experience":
        for f in fields(self):
            setattr(self, f.name, getattr(self, f.name).to(device))
        return self

    def pin_memory(self) -> "Experience":
        for f in fields(self):
            setattr(self, f.name, getattr(self, f.name).pin_memory())
        return self


class NaiveExperienceBuffer:
    """FIFO of per-sample experiences; `sample()` re-batches (right-padding sequences of different rollouts)."""

    def __init__(self, sample_batch_size: int, limit: int = 0, cpu_offload: bool = True) -> None:
        self.sample_batch_size, self.limit, self.cpu_offload = sample_batch_size, limit, cpu_offload
        self.items: List[dict] = []

    @torch.no_grad()
    def append(self, exp: Experience) -> None:
        if self.cpu_offload:
            exp = exp.to_device("cpu")
        B = exp.sequences.shape[0]
        for i in range(B):
            self.items.append({f.name: getattr(exp, f.name)[i] for f in fields(exp)})
        if self.limit > 0 and len(self.items) > self.limit:
            self.items = self.items[-self.limit:]

    def clear(self) -> None:
        self.items.clear()

    def __len__(self) -> int:
        return len(self.items)

    @staticmethod
    def _pad_stack(ts: List[torch.Tensor]) -> torch.Tensor:
        if ts[0].dim() == 0:
            return torch.stack(ts)
        n = max(t.shape[0] for t in ts)
        return torch.stack([torch.nn.functional.pad(t, (0, n - t.shape[0])) for t in ts])

    def collate(self, batch: List[dict]) -> Experience:
        return Experience(**{k: self._pad_stack([b[k] for b in batch]) for k in batch[0]})

    @torch.no_grad()
    def sample(self, device=None) -> Experience:
        exp = self.collate(random.sample(self.items, min(self.sample_batch_size, len(self.items))))
        return exp.to_device(device) if device is not None else exp


class NaiveExperienceMaker:
    def __init__(self, actor: nn.Module, critic: nn.Module, reward_model: Optional[nn.Module],
                 initial_model: nn.Module, kl_coef: float = 0.01, gamma: float = 1.0, lam: float = 0.95,
                 reward_fn: Optional[Callable] = None, pad_token_id: int = 0, eos_token_id: Optional[int] = None) -> None:
        self.actor, self.critic, self.reward_model, self.initial_model = actor, critic, reward_model, initial_model
        self.kl_coef, self.gamma, self.lam, self.reward_fn = kl_coef, gamma, lam, reward_fn
        self.pad_token_id, self.eos_token_id = pad_token_id, eos_token_id

    @staticmethod
    def gae(values: torch.Tensor, rewards: torch.Tensor, mask: torch.Tensor, gamma: float, lam: float) -> torch.Tensor:
        """Generalised advantage estimation over the action tokens (values beyond the last action are 0)."""
        A = rewards.shape[1]
        adv = torch.zeros_like(rewards)
        last = torch.zeros(rewards.shape[0], device=rewards.device, dtype=rewards.dtype)
        v = values * mask
        for t in reversed(range(A)):
            nxt = v[:, t + 1] if t + 1 < A else torch.zeros_like(last)
            delta = rewards[:, t] + gamma * nxt - v[:, t]
            last = delta + gamma * lam * last
            adv[:, t] = last
        return adv * mask

    @torch.no_grad()
    def make_experience(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                        **generate_kwargs) -> Experience:
        for m in (self.actor, self.critic, self.initial_model, self.reward_model):
            if m is not None:
                m.eval()
        prompt_len = input_ids.shape[1]
        seq = generate(self.actor, input_ids, attention_mask, eos_token_id=self.eos_token_id,
                       pad_token_id=self.pad_token_id, **generate_kwargs)
        A = seq.shape[1] - prompt_len
        gen = seq[:, prompt_len:]
        if self.eos_token_id is not None:
            # tokens up to and including the first EOS are actions
            is_eos = gen == self.eos_token_id
            after = (is_eos.long().cumsum(-1) - is_eos.long()) > 0
            action_mask = (~after).to(torch.float32)
        else:
            action_mask = torch.ones_like(gen, dtype=torch.float32)
        pmask = attention_mask if attention_mask is not None else torch.ones_like(input_ids)
        full_mask = torch.cat([pmask, action_mask.to(pmask.dtype)], dim=1)
        lp = calc_action_log_probs(get_logits(self.actor, seq, full_mask), seq, A)
        ref_lp = calc_action_log_probs(get_logits(self.initial_model, seq, full_mask), seq, A)
        values = self.critic(seq, full_mask)[:, -A - 1:-1].float()       # value BEFORE emitting each action token
        if self.reward_fn is not None:
            r = self.reward_fn(seq, prompt_len).to(lp.dtype)
        else:
            r = self.reward_model(seq, full_mask).float()
        reward, kl = compute_reward(r, self.kl_coef, lp, ref_lp, action_mask)
        adv = self.gae(values, reward, action_mask, self.gamma, self.lam)
        return Experience(seq, lp, values, r, kl, adv, full_mask, action_mask)
