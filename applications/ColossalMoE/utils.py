"""Checkpoint + metric helpers of the MoE application.  Parity: reference `applications/ColossalMoE/utils.py:1-84`
(`load_checkpoint` / `save_checkpoint` with running states, `move_to_cuda`)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


def move_to_device(batch: Dict, device) -> Dict:
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}


def save_checkpoint(save_dir: str, booster, model, optimizer, lr_scheduler, epoch: int, step: int, batch_size: int,
                    is_master: bool) -> str:
    """`save_dir/epoch-E_step-S/{modeling, optimizer, lr_scheduler, running_states.json}`: the model as a sharded
    HF-layout state dict (expert shards gathered over the expert-parallel group by the MoE checkpoint IO)."""
    path = os.path.join(save_dir, f"epoch-{epoch}_step-{step}")
    os.makedirs(os.path.join(path, "modeling"), exist_ok=True)
    booster.save_model(model, os.path.join(path, "modeling"), shard=True)
    booster.save_optimizer(optimizer, os.path.join(path, "optimizer"), shard=True)
    if lr_scheduler is not None:
        booster.save_lr_scheduler(lr_scheduler, os.path.join(path, "lr_scheduler"))
    if is_master:
        with open(os.path.join(path, "running_states.json"), "w") as f:
            json.dump({"epoch": epoch, "step": step, "sample_start_index": step * batch_size}, f, indent=2)
    return path


def load_checkpoint(load_dir: str, booster, model, optimizer=None, lr_scheduler=None) -> Tuple[int, int, int]:
    booster.load_model(model, os.path.join(load_dir, "modeling"))
    if optimizer is not None and os.path.isdir(os.path.join(load_dir, "optimizer")):
        booster.load_optimizer(optimizer, os.path.join(load_dir, "optimizer"))
    if lr_scheduler is not None and os.path.exists(os.path.join(load_dir, "lr_scheduler")):
        booster.load_lr_scheduler(lr_scheduler, os.path.join(load_dir, "lr_scheduler"))
    with open(os.path.join(load_dir, "running_states.json")) as f:
        s = json.load(f)
    return s["epoch"], s["step"], s["sample_start_index"]


class ExpertLoadMonitor:
    """Counts, per MoE layer, how many routed token slots each expert received (forward hooks on the routers: their
    output is `(weights, expert indices, logits)`).  `fractions()` = share per expert averaged over the layers, summed
    over the ranks of `group` ([num_experts], sums to 1) - what the auxiliary load-balancing loss pushes towards uniform;
    `imbalance()` = max share x num_experts (1.0 = perfectly balanced)."""

    def __init__(self, model: nn.Module) -> None:
        self.counts: Dict[int, torch.Tensor] = {}
        self.handles = []
        layer = 0
        for m in model.modules():
            router = getattr(m, "router", None)
            n = getattr(m, "num_experts", None)
            if isinstance(router, nn.Module) and n:
                self.handles.append(router.register_forward_hook(self._hook(layer, int(n))))
                layer += 1

    def _hook(self, layer: int, n: int):
        def hook(module, args, output):
            idx = output[1].detach().reshape(-1)
            c = torch.bincount(idx, minlength=n).float()
            self.counts[layer] = self.counts.get(layer, torch.zeros_like(c)) + c
        return hook

    def reset(self) -> None:
        self.counts.clear()

    def fractions(self, group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
        if not self.counts:
            return None
        c = torch.stack([self.counts[k] for k in sorted(self.counts)])
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            c = c.clone()
            dist.all_reduce(c, group=group)
        return (c / c.sum(-1, keepdim=True).clamp(min=1)).mean(0)

    def imbalance(self, group: Optional[dist.ProcessGroup] = None) -> float:
        f = self.fractions(group)
        return float(f.max() * f.numel()) if f is not None else float("nan")

    def remove(self) -> None:
        for h in self.handles:
            h.remove()
