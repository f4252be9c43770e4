"""Adafactor (factored second moments).  Parity: reference `colossalai/nn/optimizer/adafactor.py`."""
from __future__ import annotations

import math

import torch
from torch.optim import Optimizer

__all__ = ["Adafactor"]


class Adafactor(Optimizer):
    def __init__(self, params, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=None,
                 weight_decay=0.0, scale_parameter=True, relative_step=True, warmup_init=False) -> None:
        if lr is not None and relative_step:
            raise ValueError("Cannot combine manual `lr` and `relative_step=True` options")
        if warmup_init and not relative_step:
            raise ValueError("`warmup_init=True` requires `relative_step=True`")
        super().__init__(params, dict(lr=lr, eps=eps, clip_threshold=clip_threshold, decay_rate=decay_rate,
                                      beta1=beta1, weight_decay=weight_decay, scale_parameter=scale_parameter,
                                      relative_step=relative_step, warmup_init=warmup_init))

    @staticmethod
    def _get_lr(group, st):
        rel = group["lr"]
        if group["relative_step"]:
            min_step = 1e-6 * st["step"] if group["warmup_init"] else 1e-2
            rel = min(min_step, 1.0 / math.sqrt(st["step"]))
        scale = 1.0
        if group["scale_parameter"]:
            scale = max(group["eps"][1], st["RMS"])
        return scale * rel

    @staticmethod
    def _get_options(group, shape):
        return len(shape) >= 2, group["beta1"] is not None

    @staticmethod
    def _rms(t):
        return t.norm(2) / (t.numel() ** 0.5)

    @staticmethod
    def _approx_sq_grad(row, col):
        r = (row / row.mean(dim=-1, keepdim=True)).rsqrt_().unsqueeze(-1)
        c = col.unsqueeze(-2).rsqrt()
        return torch.mul(r, c)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.float()
                if grad.is_sparse:
                    raise RuntimeError("Adafactor does not support sparse gradients.")
                st = self.state[p]
                factored, use_first = self._get_options(group, grad.shape)
                if len(st) == 0:
                    st["step"] = 0
                    if use_first:
                        st["exp_avg"] = torch.zeros_like(grad)
                    if factored:
                        st["exp_avg_sq_row"] = torch.zeros(grad.shape[:-1], device=grad.device)
                        st["exp_avg_sq_col"] = torch.zeros(grad.shape[:-2] + grad.shape[-1:], device=grad.device)
                    else:
                        st["exp_avg_sq"] = torch.zeros_like(grad)
                    st["RMS"] = 0
                pf = p.float()
                st["step"] += 1
                st["RMS"] = float(self._rms(pf))
                lr = self._get_lr(group, st)
                beta2t = 1.0 - math.pow(st["step"], group["decay_rate"])
                update = grad ** 2 + group["eps"][0]
                if factored:
                    st["exp_avg_sq_row"].mul_(beta2t).add_(update.mean(dim=-1), alpha=1.0 - beta2t)
                    st["exp_avg_sq_col"].mul_(beta2t).add_(update.mean(dim=-2), alpha=1.0 - beta2t)
                    update = self._approx_sq_grad(st["exp_avg_sq_row"], st["exp_avg_sq_col"]).mul_(grad)
                else:
                    st["exp_avg_sq"].mul_(beta2t).add_(update, alpha=1.0 - beta2t)
                    update = st["exp_avg_sq"].rsqrt().mul_(grad)
                update.div_((self._rms(update) / group["clip_threshold"]).clamp_(min=1.0))
                update.mul_(lr)
                if use_first:
                    st["exp_avg"].mul_(group["beta1"]).add_(update, alpha=1 - group["beta1"])
                    update = st["exp_avg"]
                if group["weight_decay"] != 0:
                    pf = pf - group["weight_decay"] * lr * pf
                p.copy_(pf - update)
        return loss
