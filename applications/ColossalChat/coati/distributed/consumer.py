"""GRPO / DAPO consumer: receives rollouts, scores them with a verifiable reward, computes group-relative advantages
and updates the policy through a Booster; publishes fresh weights for the producers.
Parity: reference `coati/distributed/{consumer.py:1-400 (BaseConsumer), grpo_consumer.py:1-600 (GRPOConsumer)}`."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from ..models import PolicyLoss, calc_action_log_probs, get_logits
from ..trainer.grpo import group_advantages


class GRPOConsumer:
    def __init__(self, policy: nn.Module, optimizer, reward_fn: Callable, booster=None,
                 reference: Optional[nn.Module] = None, num_generations: int = 4, beta: float = 0.0,
                 clip_eps_low: float = 0.2, clip_eps_high: Optional[float] = None, loss_variation: str = "sample_level",
                 filter_uniform_groups: bool = False, eos_token_id: Optional[int] = None,
                 minibatch_size: Optional[int] = None) -> None:
        self.policy, self.optimizer, self.booster, self.reference = policy, optimizer, booster, reference
        self.reward_fn, self.G, self.eos_token_id = reward_fn, num_generations, eos_token_id
        self.filter_uniform, self.minibatch_size = filter_uniform_groups, minibatch_size
        self.loss_fn = PolicyLoss(clip_eps_low, clip_eps_high, beta=beta if reference is not None else 0.0,
                                  loss_variation=loss_variation)
        self.version = 0
        self.history: List[Dict[str, float]] = []

    def _action_mask(self, gen: torch.Tensor) -> torch.Tensor:
        if self.eos_token_id is None:
            return torch.ones_like(gen, dtype=torch.float32)
        is_eos = gen == self.eos_token_id
        return (~((is_eos.long().cumsum(-1) - is_eos.long()) > 0)).float()

    def step(self, rollout: Dict) -> Dict[str, float]:
        seq, P = rollout["sequences"], rollout["prompt_len"]
        dev = next(self.policy.parameters()).device
        seq = seq.to(dev)
        action_mask = self._action_mask(seq[:, P:])
        full_mask = torch.cat([rollout["attention_mask"].to(dev), action_mask.to(rollout["attention_mask"].dtype)], 1)
        extra = {k: v for k, v in rollout.items() if isinstance(v, list)}
        rewards = self.reward_fn(seq, P, **extra).float().to(dev)
        adv = group_advantages(rewards, self.G)
        keep = torch.ones_like(rewards, dtype=torch.bool)
        if self.filter_uniform:
            keep = (rewards.view(-1, self.G).std(1, unbiased=False) > 0).repeat_interleave(self.G)
        metrics = {"reward": float(rewards.mean()), "kept": float(keep.float().mean()),
                   "response_len": float(action_mask.sum(-1).mean()), "staleness": float(self.version - rollout.get("model_version", self.version))}
        if int(keep.sum()) == 0:
            self.history.append(metrics)
            return metrics
        seq, full_mask, action_mask, adv = seq[keep], full_mask[keep], action_mask[keep], adv[keep]
        A = action_mask.shape[1]
        self.policy.train()
        with torch.no_grad():
            old_lp = calc_action_log_probs(get_logits(self.policy, seq, full_mask), seq, A)
            ref_lp = calc_action_log_probs(get_logits(self.reference, seq, full_mask), seq, A) \
                if self.reference is not None else None
        mb = self.minibatch_size or seq.shape[0]
        total = 0.0
        for s in range(0, seq.shape[0], mb):
            sl = slice(s, s + mb)
            lp = calc_action_log_probs(get_logits(self.policy, seq[sl], full_mask[sl]), seq[sl], A)
            kl = None
            if ref_lp is not None:
                d = ref_lp[sl] - lp
                kl = d.exp() - d - 1
            loss, skipped, _ = self.loss_fn(lp, old_lp[sl], adv[sl], action_mask[sl], kl)
            if skipped:
                continue
            loss = loss * (min(mb, seq.shape[0] - s) / seq.shape[0])
            if self.booster is not None:
                self.booster.backward(loss, self.optimizer)
            else:
                loss.backward()
            total += float(loss.detach())
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.version += 1
        metrics["loss"] = total
        self.history.append(metrics)
        return metrics

    def save_checkpoint(self, directory: str, shard: bool = False) -> str:
        """`<directory>/step_<version>/{model, optimizer, state.json}` through the booster when there is one (sharded,
        parallel-aware) or `torch.save` otherwise; returns the path (reference `BaseConsumer` saves every
        `save_interval` updates, `consumer.py:330-360`)."""
        import json
        import os

        path = os.path.join(directory, f"step_{self.version}")
        os.makedirs(path, exist_ok=True)
        if self.booster is not None:
            self.booster.save_model(self.policy, os.path.join(path, "model"), shard=shard)
            self.booster.save_optimizer(self.optimizer, os.path.join(path, "optimizer"), shard=shard)
        else:
            torch.save(self.policy.state_dict(), os.path.join(path, "model.pt"))
            torch.save(self.optimizer.state_dict(), os.path.join(path, "optimizer.pt"))
        import torch.distributed as dist

        writer = getattr(self, "checkpoint_writer", None)      # set by `launch_distributed` (first consumer rank)
        if writer is None:
            writer = not dist.is_initialized() or self.booster is None or dist.get_rank() == 0
        if writer:
            with open(os.path.join(path, "state.json"), "w") as f:
                json.dump({"version": self.version, "last": self.history[-1] if self.history else {}}, f)
        return path

    def state_dict_for_producers(self) -> Dict[str, torch.Tensor]:
        from colossalai_b200.interface import ModelWrapper

        m = self.policy.unwrap() if isinstance(self.policy, ModelWrapper) else self.policy
        return {k: v.detach() for k, v in m.state_dict().items()}
