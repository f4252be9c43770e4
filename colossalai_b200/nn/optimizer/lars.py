"""LARS.  Parity: reference `colossalai/nn/optimizer/lars.py`."""
from __future__ import annotations

from typing import Iterable

import torch
from torch.optim import Optimizer

__all__ = ["Lars"]


class Lars(Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, momentum=0, eeta=1e-3, weight_decay=0,
                 epsilon=0.0) -> None:
        if not isinstance(lr, float) or lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if eeta <= 0 or eeta > 1:
            raise ValueError(f"Invalid eeta value: {eeta}")
        if epsilon < 0:
            raise ValueError(f"Invalid epsilon value: {epsilon}")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, eeta=eeta, epsilon=epsilon,
                                      lars=True))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            wd, mom, eeta, lr, eps = group["weight_decay"], group["momentum"], group["eeta"], group["lr"], group["epsilon"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                d = p.grad.float()
                pf = p.float()
                scaled_lr = lr
                if group["lars"]:
                    w_norm, g_norm = pf.norm(), d.norm()
                    trust = torch.where((w_norm > 0) & (g_norm > 0), eeta * w_norm / (g_norm + wd * w_norm + eps),
                                        torch.ones_like(w_norm))
                    scaled_lr = lr * trust
                    if wd != 0:
                        d = d + wd * pf
                if mom != 0:
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        buf = st["momentum_buffer"] = d.clone()
                    else:
                        buf = st["momentum_buffer"]
                        buf.mul_(mom).add_(d)
                    d = buf
                p.copy_(pf - scaled_lr * d)
        return loss
