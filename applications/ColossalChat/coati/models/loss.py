"""Objectives of the alignment trainers, written from the papers' formulas.
Parity: reference `coati/models/loss.py:14-282` (`GPTLMLoss`, `PolicyLoss`, `ValueLoss`, `DpoLoss`, `LogSigLoss`,
`LogExpLoss`, `OddsRatioLoss`, `KTOLoss`) + `coati/distributed/loss.py` (token-level GRPO / DAPO policy loss)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .utils import masked_mean


class GPTLMLoss(nn.Module):
    """Next-token cross entropy (labels == -100 ignored)."""

    def forward(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1),
                               ignore_index=-100)


class PolicyLoss(nn.Module):
    """PPO clipped surrogate.  `clip_eps_low/high` differ for DAPO ("clip higher"); `beta` > 0 adds the k3 KL
    estimator against the reference policy (GRPO); `loss_variation` = "sample_level" averages per sequence first,
    "token_level" averages over all action tokens of the batch (DAPO)."""

    def __init__(self, clip_eps_low: float = 0.2, clip_eps_high: Optional[float] = None, skip_threshold: float = 20.0,
                 beta: float = 0.0, loss_variation: str = "sample_level") -> None:
        super().__init__()
        self.eps_low, self.eps_high = clip_eps_low, clip_eps_high if clip_eps_high is not None else clip_eps_low
        self.skip_threshold, self.beta, self.loss_variation = skip_threshold, beta, loss_variation

    def forward(self, log_probs: torch.Tensor, old_log_probs: torch.Tensor, advantages: torch.Tensor,
                action_mask: Optional[torch.Tensor] = None, per_token_kl: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, bool, torch.Tensor]:
        ratio_ = ((log_probs - old_log_probs) * (action_mask if action_mask is not None else 1.0)).exp()
        # a wildly off-policy batch (stale rollouts, numerical blow-up) is skipped instead of poisoning the update
        if float(ratio_.detach().max()) > self.skip_threshold:
            return log_probs.sum() * 0.0, True, ratio_.detach().max()
        ratio = ratio_.clamp(0.0, 10.0)
        if advantages.dim() == 1:
            advantages = advantages[:, None]
        surr = torch.min(ratio * advantages, ratio.clamp(1 - self.eps_low, 1 + self.eps_high) * advantages)
        loss = -surr
        if self.beta > 0 and per_token_kl is not None:
            loss = loss + self.beta * per_token_kl
        if self.loss_variation == "token_level" and action_mask is not None:
            m = action_mask.to(loss.dtype)
            loss = (loss * m).sum() / m.sum().clamp(min=1)
        else:
            loss = masked_mean(loss, action_mask, dim=-1).mean()
        return loss, False, ratio_.detach().max()


class ValueLoss(nn.Module):
    """Clipped value loss of PPO."""

    def __init__(self, clip_eps: float = 0.2) -> None:
        super().__init__()
        self.clip_eps = clip_eps

    def forward(self, values: torch.Tensor, old_values: torch.Tensor, returns: torch.Tensor,
                action_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        clipped = old_values + (values - old_values).clamp(-self.clip_eps, self.clip_eps)
        loss = torch.max((clipped - returns) ** 2, (values - returns) ** 2)
        return 0.5 * masked_mean(loss, action_mask, dim=-1).mean()


class DpoLoss(nn.Module):
    """-log sigmoid(beta * [(pi_c - ref_c) - (pi_r - ref_r)] - gamma); without a reference model and with `gamma` > 0
    and length-normalised log-probs this is SimPO."""

    def __init__(self, beta: float = 0.1, gamma: float = 0.0) -> None:
        super().__init__()
        self.beta, self.gamma = beta, gamma

    def forward(self, logprob_actor_chosen: torch.Tensor, logprob_actor_reject: torch.Tensor,
                logprob_ref_chosen: Optional[torch.Tensor], logprob_ref_reject: Optional[torch.Tensor],
                chosen_mask: torch.Tensor, reject_mask: torch.Tensor):
        pc = (logprob_actor_chosen * chosen_mask).sum(-1)
        pr = (logprob_actor_reject * reject_mask).sum(-1)
        if logprob_ref_chosen is not None and logprob_ref_reject is not None:
            rc = (logprob_ref_chosen * chosen_mask).sum(-1)
            rr = (logprob_ref_reject * reject_mask).sum(-1)
        else:
            rc = rr = torch.zeros_like(pc)
        logits = (pc - rc) - (pr - rr) - self.gamma / self.beta
        losses = -F.logsigmoid(self.beta * logits)
        chosen_rewards = self.beta * (pc - rc).detach()
        rejected_rewards = self.beta * (pr - rr).detach()
        return losses.mean(), chosen_rewards, rejected_rewards


class LogSigLoss(nn.Module):
    """Pairwise ranking loss of the reward model (InstructGPT): -log sigmoid(r_c - r_r)."""

    def forward(self, chosen_reward: torch.Tensor, reject_reward: torch.Tensor) -> torch.Tensor:
        return -F.logsigmoid(chosen_reward - reject_reward).mean()


class LogExpLoss(nn.Module):
    """log(1 + exp(r_r - r_c)) (the same ranking loss in soft-plus form)."""

    def forward(self, chosen_reward: torch.Tensor, reject_reward: torch.Tensor) -> torch.Tensor:
        return F.softplus(reject_reward - chosen_reward).mean()


class OddsRatioLoss(nn.Module):
    """ORPO penalty: -log sigmoid(log odds(chosen) - log odds(rejected)) on length-normalised log-probs."""

    def forward(self, chosen_logp: torch.Tensor, reject_logp: torch.Tensor, chosen_mask: torch.Tensor,
                reject_mask: torch.Tensor):
        c = (chosen_logp * chosen_mask).sum(-1) / chosen_mask.sum(-1).clamp(min=1)
        r = (reject_logp * reject_mask).sum(-1) / reject_mask.sum(-1).clamp(min=1)
        c, r = c.float(), r.float()
        log_odds = (c - r) - (torch.log1p(-c.exp().clamp(max=1 - 1e-6)) - torch.log1p(-r.exp().clamp(max=1 - 1e-6)))
        return -F.logsigmoid(log_odds).mean(), log_odds.detach()


class KTOLoss(nn.Module):
    """Kahneman-Tversky optimisation: desirable samples push sigmoid(beta (r - z0)) up, undesirable ones push
    sigmoid(beta (z0 - r)) up, with z0 = batch estimate of KL(pi || ref) (clamped at 0, no gradient)."""

    def __init__(self, beta: float = 0.1, desirable_weight: float = 1.0, undesirable_weight: float = 1.0) -> None:
        super().__init__()
        self.beta, self.w_d, self.w_u = beta, desirable_weight, undesirable_weight

    def forward(self, chosen_logps: torch.Tensor, rejected_logps: torch.Tensor, kl_logps: torch.Tensor,
                ref_chosen_logps: torch.Tensor, ref_rejected_logps: torch.Tensor, ref_kl_logps: torch.Tensor):
        kl = (kl_logps - ref_kl_logps).mean().detach().clamp(min=0)
        losses = []
        chosen_rewards = rejected_rewards = torch.zeros(0, device=kl.device)
        if chosen_logps.numel():
            lr = chosen_logps - ref_chosen_logps
            losses.append(self.w_d * (1 - torch.sigmoid(self.beta * (lr - kl))))
            chosen_rewards = self.beta * lr.detach()
        if rejected_logps.numel():
            lr = rejected_logps - ref_rejected_logps
            losses.append(self.w_u * (1 - torch.sigmoid(self.beta * (kl - lr))))
            rejected_rewards = self.beta * lr.detach()
        return torch.cat(losses).mean(), chosen_rewards, rejected_rewards, kl
