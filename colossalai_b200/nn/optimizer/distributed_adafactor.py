"""DistributedAdaFactor: Adafactor on TP-sharded (row / column) parameters — row/col statistics of the factored second
moment are all-reduced over the TP group along the sharded dimension so the update equals the un-sharded one.
Parity: reference `colossalai/nn/optimizer/distributed_adafactor.py`."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.distributed as dist

from ...interface.optimizer import DistributedOptim

__all__ = ["DistributedAdaFactor"]


class DistributedAdaFactor(DistributedOptim):
    def __init__(self, params, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=None,
                 weight_decay=0.0, scale_parameter=True, relative_step=True, warmup_init=False) -> None:
        if lr is not None and relative_step:
            raise ValueError("Cannot combine manual `lr` and `relative_step=True` options")
        super().__init__(params, dict(lr=lr, eps=eps, clip_threshold=clip_threshold, decay_rate=decay_rate,
                                      beta1=beta1, weight_decay=weight_decay, scale_parameter=scale_parameter,
                                      relative_step=relative_step, warmup_init=warmup_init))
        self._post_cast()

    def _post_cast(self) -> None:
        for a, v in (("tp_size", 1), ("tp_group", None), ("dp_size", 1), ("dp_group", None), ("is_zero", False),
                     ("shard_to_working_param", {})):
            if not hasattr(self, a):
                setattr(self, a, v)

    def setup_distributed(self, tp_group=None, dp_group=None, shard_to_working_param: Optional[Dict] = {},
                          padding_map=None, is_zero: Optional[bool] = False) -> None:
        self.tp_group, self.dp_group = tp_group, dp_group
        self.tp_size = dist.get_world_size(tp_group) if tp_group is not None else 1
        self.dp_size = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.shard_to_working_param = shard_to_working_param or {}
        self.is_zero = bool(is_zero)

    def _shard_dim(self, p) -> Optional[int]:
        wp = self.shard_to_working_param.get(id(p), p)
        sh = getattr(wp, "dist_shard", None)
        if sh is None and hasattr(wp, "tp_shard_dim"):       # fused (customised) column / row parallel weights
            sh = (wp.tp_shard_dim, None)
        if sh is not None and self.tp_size > 1:
            return sh[0]
        return None

    def _global_mean_sq(self, t: torch.Tensor, sharded: bool) -> torch.Tensor:
        s = torch.stack([t.pow(2).sum(), torch.tensor(float(t.numel()), device=t.device)])
        if sharded:
            dist.all_reduce(s, group=self.tp_group)
        return (s[0] / s[1]).sqrt()

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
                sd = self._shard_dim(p)
                if len(st) == 0:
                    st["step"] = 0
                    if group["beta1"] is not None:
                        st["exp_avg"] = torch.zeros_like(grad)
                    if factored:
                        st["exp_avg_sq_row"] = torch.zeros(grad.shape[:-1], device=grad.device)
                        st["exp_avg_sq_col"] = torch.zeros(grad.shape[:-2] + grad.shape[-1:], device=grad.device)
                    else:
                        st["exp_avg_sq"] = torch.zeros_like(grad)
                pf = p.float()
                st["step"] += 1
                rms_p = float(self._global_mean_sq(pf, sd is not None))
                rel = group["lr"]
                if group["relative_step"]:
                    min_step = 1e-6 * st["step"] if group["warmup_init"] else 1e-2
                    rel = min(min_step, 1.0 / math.sqrt(st["step"]))
                lr = (max(group["eps"][1], rms_p) if group["scale_parameter"] else 1.0) * rel
                beta2t = 1.0 - math.pow(st["step"], group["decay_rate"])
                update = grad ** 2 + group["eps"][0]
                if factored:
                    row_mean = update.mean(dim=-1)        # over columns
                    col_mean = update.mean(dim=-2)        # over rows
                    nd = grad.dim()
                    if sd is not None and sd == nd - 1:   # columns sharded: row means need the global column average
                        dist.all_reduce(row_mean, group=self.tp_group)
                        row_mean /= self.tp_size
                    if sd is not None and sd == nd - 2:   # rows sharded
                        dist.all_reduce(col_mean, group=self.tp_group)
                        col_mean /= self.tp_size
                    st["exp_avg_sq_row"].mul_(beta2t).add_(row_mean, alpha=1.0 - beta2t)
                    st["exp_avg_sq_col"].mul_(beta2t).add_(col_mean, alpha=1.0 - beta2t)
                    row = st["exp_avg_sq_row"]
                    row_avg = row.mean(dim=-1, keepdim=True)
                    if sd is not None and sd == nd - 2:   # the mean over rows spans ranks
                        dist.all_reduce(row_avg, group=self.tp_group)
                        row_avg = row_avg / self.tp_size
                    r = (row / row_avg).rsqrt().unsqueeze(-1)
                    c = st["exp_avg_sq_col"].unsqueeze(-2).rsqrt()
                    update = torch.mul(r, c).mul_(grad)
                else:
                    st["exp_avg_sq"].mul_(beta2t).add_(update, alpha=1.0 - beta2t)
                    update = st["exp_avg_sq"].rsqrt().mul_(grad)
                rms_u = self._global_mean_sq(update, sd is not None)
                update.div_((rms_u / group["clip_threshold"]).clamp_(min=1.0))
                update.mul_(lr)
                if group["beta1"] is not None:
                    st["exp_avg"].mul_(group["beta1"]).add_(update, alpha=1 - group["beta1"])
                    update = st["exp_avg"]
                if group["weight_decay"] != 0:
                    pf = pf - group["weight_decay"] * lr * pf
                p.copy_(pf - update)
        return loss
