"""GRPO (critic-free, group-relative advantages) with the DAPO switches: asymmetric clipping, token-level loss,
dynamic filtering of groups whose rewards are all equal, over-length soft punishment.
Parity: reference `coati/trainer/grpo.py:1-386` and `coati/distributed/grpo_consumer.py:1-600`."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from ..models import PolicyLoss, calc_action_log_probs, generate, get_logits
from .base import OLTrainer, all_reduce_mean


def group_advantages(rewards: torch.Tensor, num_generations: int, eps: float = 1e-4) -> torch.Tensor:
    """(r - mean_group) / (std_group + eps) for `rewards` laid out as [num_prompts * num_generations]."""
    g = rewards.view(-1, num_generations)
    return ((g - g.mean(1, keepdim=True)) / (g.std(1, keepdim=True, unbiased=False) + eps)).reshape(-1)


class GRPOTrainer(OLTrainer):
    def __init__(self, actor_booster, actor: nn.Module, initial_model: Optional[nn.Module], actor_optim,
                 reward_fn: Callable, actor_lr_scheduler=None, num_generations: int = 4, beta: float = 0.01,
                 clip_eps_low: float = 0.2, clip_eps_high: Optional[float] = None,
                 loss_variation: str = "sample_level", filter_uniform_groups: bool = False,
                 max_response_len: Optional[int] = None, overlength_cache: int = 0, pad_token_id: int = 0,
                 eos_token_id: Optional[int] = None, generate_kwargs: Optional[Dict] = None, device=None) -> None:
        super().__init__(actor_booster, actor_optim, actor_lr_scheduler, 1, device)
        self.actor, self.initial_model, self.reward_fn = actor, initial_model, reward_fn
        if initial_model is not None:
            initial_model.eval()
            for p in initial_model.parameters():
                p.requires_grad_(False)
        self.G, self.filter_uniform = num_generations, filter_uniform_groups
        self.loss_fn = PolicyLoss(clip_eps_low, clip_eps_high, beta=beta if initial_model is not None else 0.0,
                                  loss_variation=loss_variation)
        self.max_response_len, self.overlength_cache = max_response_len, overlength_cache
        self.pad_token_id, self.eos_token_id = pad_token_id, eos_token_id
        self.generate_kwargs = generate_kwargs or {}
        self.rollouts: List[Dict[str, torch.Tensor]] = []

    @torch.no_grad()
    def _collect(self, prompts) -> None:
        ids = prompts["input_ids"].repeat_interleave(self.G, dim=0)
        am = prompts.get("attention_mask")
        am = am.repeat_interleave(self.G, dim=0) if am is not None else torch.ones_like(ids)
        P = ids.shape[1]
        seq = generate(self.actor, ids, am, eos_token_id=self.eos_token_id, pad_token_id=self.pad_token_id,
                       **self.generate_kwargs)
        gen = seq[:, P:]
        if self.eos_token_id is not None:
            is_eos = gen == self.eos_token_id
            action_mask = (~((is_eos.long().cumsum(-1) - is_eos.long()) > 0)).float()
        else:
            action_mask = torch.ones_like(gen, dtype=torch.float32)
        full_mask = torch.cat([am, action_mask.to(am.dtype)], 1)
        extra = {k: [x for x in v for _ in range(self.G)] for k, v in prompts.items() if isinstance(v, list)}
        rewards = self.reward_fn(seq, P, **extra).float().to(seq.device)
        if self.max_response_len is not None and self.overlength_cache > 0:       # DAPO soft over-length punishment
            L = action_mask.sum(-1)
            start = self.max_response_len - self.overlength_cache
            rewards = rewards - ((L - start).clamp(min=0) / self.overlength_cache).clamp(max=1.0)
        adv = group_advantages(rewards, self.G)
        keep = torch.ones_like(rewards, dtype=torch.bool)
        if self.filter_uniform:                                                    # DAPO dynamic sampling
            g = rewards.view(-1, self.G)
            keep = (g.std(1, unbiased=False) > 0).repeat_interleave(self.G)
        A = gen.shape[1]
        old_lp = calc_action_log_probs(get_logits(self.actor, seq, full_mask), seq, A)
        ref_lp = calc_action_log_probs(get_logits(self.initial_model, seq, full_mask), seq, A) \
            if self.initial_model is not None else None
        self.rollouts.append({"seq": seq[keep], "mask": full_mask[keep], "action_mask": action_mask[keep],
                              "adv": adv[keep], "old_lp": old_lp[keep], "reward": rewards,
                              "ref_lp": None if ref_lp is None else ref_lp[keep]})

    def _update(self) -> Dict[str, float]:
        self.actor.train()
        stats = {"loss": 0.0, "reward": 0.0, "kl": 0.0, "n": 0}
        for ro in self.rollouts:
            stats["reward"] += float(ro["reward"].mean())
            if ro["seq"].shape[0] == 0:
                continue
            A = ro["action_mask"].shape[1]
            lp = calc_action_log_probs(get_logits(self.actor, ro["seq"], ro["mask"]), ro["seq"], A)
            kl = None
            if ro["ref_lp"] is not None:            # k3 estimator: exp(ref - pi) - (ref - pi) - 1 >= 0
                d = ro["ref_lp"] - lp
                kl = d.exp() - d - 1
            loss, skipped, _ = self.loss_fn(lp, ro["old_lp"], ro["adv"], ro["action_mask"], kl)
            if not skipped:
                if self.booster is not None:
                    self.booster.backward(loss, self.optimizer)
                else:
                    loss.backward()
                self.optimizer.step()
                if self.lr_scheduler is not None:
                    self.lr_scheduler.step()
            self.optimizer.zero_grad()
            stats["loss"] += float(loss.detach())
            stats["kl"] += float((kl * ro["action_mask"]).sum() / ro["action_mask"].sum().clamp(min=1)) if kl is not None else 0.0
            stats["n"] += 1
        n = max(1, len(self.rollouts))
        out = {"loss": stats["loss"] / max(1, stats["n"]), "reward": stats["reward"] / n, "kl": stats["kl"] / max(1, stats["n"])}
        return {k: float(all_reduce_mean(torch.tensor(v))) for k, v in out.items()}

    def _after_episode(self) -> None:
        self.rollouts.clear()
