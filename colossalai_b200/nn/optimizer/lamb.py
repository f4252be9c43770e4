"""Lamb (pure PyTorch).  Parity: reference `colossalai/nn/optimizer/lamb.py`."""
from __future__ import annotations

import torch
from torch.optim import Optimizer

__all__ = ["Lamb"]


class Lamb(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, adam=False,
                 bias_correction=False) -> None:
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      bias_correction=bias_correction))
        self.adam = adam

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.float()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                b1, b2 = group["betas"]
                st["step"] += 1
                st["exp_avg"].mul_(b1).add_(grad, alpha=1 - b1)
                st["exp_avg_sq"].mul_(b2).addcmul_(grad, grad, value=1 - b2)
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if group["bias_correction"]:
                    m = m / (1 - b1 ** st["step"])
                    v = v / (1 - b2 ** st["step"])
                update = m / (v.sqrt() + group["eps"])
                pf = p.float()
                if group["weight_decay"] != 0:
                    update = update + group["weight_decay"] * pf
                w_norm, u_norm = pf.norm(), update.norm()
                trust = torch.where((w_norm > 0) & (u_norm > 0), w_norm / u_norm, torch.ones_like(w_norm))
                if self.adam:
                    trust = torch.ones_like(trust)
                p.copy_(pf - group["lr"] * trust * update)
        return loss
