"""CAME (confidence-guided adaptive memory-efficient optimizer).  Parity: reference `colossalai/nn/optimizer/came.py`."""
from __future__ import annotations

import torch
from torch.optim import Optimizer

__all__ = ["CAME"]


class CAME(Optimizer):
    def __init__(self, params, lr=None, eps=(1e-30, 1e-16), clip_threshold=1.0, betas=(0.9, 0.999, 0.9999),
                 weight_decay=0.0) -> None:
        assert lr is not None and lr > 0.0
        assert all(0.0 <= b <= 1.0 for b in betas)
        super().__init__(params, dict(lr=lr, eps=eps, clip_threshold=clip_threshold, betas=betas,
                                      weight_decay=weight_decay))

    @property
    def supports_memory_efficient_fp16(self):
        return True

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
                st = self.state[p]
                factored = grad.dim() >= 2
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(grad)
                    if factored:
                        st["exp_avg_sq_row"] = torch.zeros(grad.shape[:-1], device=grad.device)
                        st["exp_avg_sq_col"] = torch.zeros(grad.shape[:-2] + grad.shape[-1:], device=grad.device)
                        st["exp_avg_res_row"] = torch.zeros(grad.shape[:-1], device=grad.device)
                        st["exp_avg_res_col"] = torch.zeros(grad.shape[:-2] + grad.shape[-1:], device=grad.device)
                    else:
                        st["exp_avg_sq"] = torch.zeros_like(grad)
                st["step"] += 1
                b1, b2, b3 = group["betas"]
                update = grad ** 2 + group["eps"][0]
                if factored:
                    st["exp_avg_sq_row"].mul_(b2).add_(update.mean(dim=-1), alpha=1.0 - b2)
                    st["exp_avg_sq_col"].mul_(b2).add_(update.mean(dim=-2), alpha=1.0 - b2)
                    update = self._approx_sq_grad(st["exp_avg_sq_row"], st["exp_avg_sq_col"]).mul_(grad)
                else:
                    st["exp_avg_sq"].mul_(b2).add_(update, alpha=1.0 - b2)
                    update = st["exp_avg_sq"].rsqrt().mul_(grad)
                update.div_((self._rms(update) / group["clip_threshold"]).clamp_(min=1.0))
                st["exp_avg"].mul_(b1).add_(update, alpha=1 - b1)
                if factored:
                    res = (update - st["exp_avg"]) ** 2 + group["eps"][1]
                    st["exp_avg_res_row"].mul_(b3).add_(res.mean(dim=-1), alpha=1.0 - b3)
                    st["exp_avg_res_col"].mul_(b3).add_(res.mean(dim=-2), alpha=1.0 - b3)
                    update = self._approx_sq_grad(st["exp_avg_res_row"], st["exp_avg_res_col"]).mul_(st["exp_avg"])
                else:
                    update = st["exp_avg"].clone()
                pf = p.float()
                if group["weight_decay"] != 0:
                    pf = pf - group["weight_decay"] * group["lr"] * pf
                p.copy_(pf - group["lr"] * update)
        return loss
