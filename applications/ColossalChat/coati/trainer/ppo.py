"""PPO with a learned critic, a frozen reference policy and a reward model (or a programmatic reward).
Parity: reference `coati/trainer/ppo.py:1-412`."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

from ..experience import NaiveExperienceBuffer, NaiveExperienceMaker
from ..models import GPTLMLoss, PolicyLoss, ValueLoss, calc_action_log_probs, get_logits, masked_whiten
from .base import OLTrainer, all_reduce_mean


class PPOTrainer(OLTrainer):
    def __init__(self, actor_booster, critic_booster, actor: nn.Module, critic: nn.Module,
                 reward_model: Optional[nn.Module], initial_model: nn.Module, actor_optim, critic_optim,
                 actor_lr_scheduler=None, critic_lr_scheduler=None, kl_coef: float = 0.02, ptx_coef: float = 0.0,
                 train_batch_size: int = 8, buffer_limit: int = 0, eps_clip: float = 0.2, vf_coef: float = 1.0,
                 value_clip: float = 0.2, gamma: float = 1.0, lam: float = 0.95, reward_fn: Optional[Callable] = None,
                 pad_token_id: int = 0, eos_token_id: Optional[int] = None, whiten_advantages: bool = True,
                 generate_kwargs: Optional[Dict] = None, device=None) -> None:
        super().__init__(actor_booster, actor_optim, actor_lr_scheduler, 1, device)
        self.actor, self.critic = actor, critic
        self.critic_booster, self.critic_optim, self.critic_lr_scheduler = critic_booster, critic_optim, critic_lr_scheduler
        for m in (reward_model, initial_model):
            if m is not None:
                m.eval()
                for p in m.parameters():
                    p.requires_grad_(False)
        self.maker = NaiveExperienceMaker(actor, critic, reward_model, initial_model, kl_coef, gamma, lam, reward_fn,
                                          pad_token_id, eos_token_id)
        self.buffer = NaiveExperienceBuffer(train_batch_size, buffer_limit)
        self.actor_loss, self.critic_loss, self.ptx_loss = PolicyLoss(eps_clip), ValueLoss(value_clip), GPTLMLoss()
        self.vf_coef, self.ptx_coef, self.whiten = vf_coef, ptx_coef, whiten_advantages
        self.generate_kwargs = generate_kwargs or {}
        self.pretrain_batch: Optional[Dict[str, torch.Tensor]] = None

    def _collect(self, prompts) -> None:
        exp = self.maker.make_experience(prompts["input_ids"], prompts.get("attention_mask"), **self.generate_kwargs)
        self.buffer.append(exp)

    def _update(self) -> Dict[str, float]:
        self.actor.train()
        self.critic.train()
        exp = self.buffer.sample(self.device)
        A = exp.action_mask.shape[1]
        adv = masked_whiten(exp.advantages, exp.action_mask) if self.whiten else exp.advantages
        lp = calc_action_log_probs(get_logits(self.actor, exp.sequences, exp.attention_mask), exp.sequences, A)
        a_loss, skipped, max_ratio = self.actor_loss(lp, exp.action_log_probs, adv, exp.action_mask)
        if self.ptx_coef > 0 and self.pretrain_batch is not None:
            pb = self._to_device(self.pretrain_batch)
            a_loss = a_loss + self.ptx_coef * self.ptx_loss(get_logits(self.actor, pb["input_ids"],
                                                                       pb.get("attention_mask")), pb["labels"])
        if not skipped:
            if self.booster is not None:
                self.booster.backward(a_loss, self.optimizer)
            else:
                a_loss.backward()
            self.optimizer.step()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
        self.optimizer.zero_grad()
        values = self.critic(exp.sequences, exp.attention_mask)[:, -A - 1:-1].float()
        returns = exp.advantages + exp.values
        c_loss = self.vf_coef * self.critic_loss(values, exp.values, returns, exp.action_mask)
        if self.critic_booster is not None:
            self.critic_booster.backward(c_loss, self.critic_optim)
        else:
            c_loss.backward()
        self.critic_optim.step()
        self.critic_optim.zero_grad()
        if self.critic_lr_scheduler is not None:
            self.critic_lr_scheduler.step()
        return {"actor_loss": float(all_reduce_mean(a_loss.detach())), "critic_loss": float(all_reduce_mean(c_loss.detach())),
                "reward": float(all_reduce_mean(exp.reward.mean())), "kl": float(all_reduce_mean(exp.kl.mean())),
                "max_ratio": float(max_ratio), "skipped": float(skipped)}

    def _after_episode(self) -> None:
        self.buffer.clear()
