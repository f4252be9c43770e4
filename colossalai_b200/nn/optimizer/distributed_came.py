"""DistributedCAME: CAME with TP-aware row/column statistics (same reduction scheme as DistributedAdaFactor).
Parity: reference `colossalai/nn/optimizer/distributed_came.py`."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from ...interface.optimizer import DistributedOptim

__all__ = ["DistributedCAME"]


class DistributedCAME(DistributedOptim):
    def __init__(self, params, lr=None, eps=(1e-30, 1e-16), clip_threshold=1.0, betas=(0.9, 0.999, 0.9999),
                 weight_decay=0.0) -> None:
        assert lr is not None and lr > 0.0
        super().__init__(params, dict(lr=lr, eps=eps, clip_threshold=clip_threshold, betas=betas,
                                      weight_decay=weight_decay))
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

    def _shard_dim(self, p):
        wp = self.shard_to_working_param.get(id(p), p)
        sh = getattr(wp, "dist_shard", None)
        if sh is None and hasattr(wp, "tp_shard_dim"):       # fused (customised) column / row parallel weights
            sh = (wp.tp_shard_dim, None)
        return sh[0] if (sh is not None and self.tp_size > 1) else None

    def _rms(self, t, sharded):
        s = torch.stack([t.pow(2).sum(), torch.tensor(float(t.numel()), device=t.device)])
        if sharded:
            dist.all_reduce(s, group=self.tp_group)
        return (s[0] / s[1]).sqrt()

    def _factored(self, stat_row, stat_col, x, beta, sd, nd):
        row_mean, col_mean = x.mean(dim=-1), x.mean(dim=-2)
        if sd is not None and sd == nd - 1:
            dist.all_reduce(row_mean, group=self.tp_group)
            row_mean /= self.tp_size
        if sd is not None and sd == nd - 2:
            dist.all_reduce(col_mean, group=self.tp_group)
            col_mean /= self.tp_size
        stat_row.mul_(beta).add_(row_mean, alpha=1.0 - beta)
        stat_col.mul_(beta).add_(col_mean, alpha=1.0 - beta)
        row_avg = stat_row.mean(dim=-1, keepdim=True)
        if sd is not None and sd == nd - 2:
            dist.all_reduce(row_avg, group=self.tp_group)
            row_avg = row_avg / self.tp_size
        return torch.mul((stat_row / row_avg).rsqrt().unsqueeze(-1), stat_col.unsqueeze(-2).rsqrt())

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2, b3 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.float()
                st = self.state[p]
                factored = grad.dim() >= 2
                sd, nd = self._shard_dim(p), grad.dim()
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(grad)
                    if factored:
                        for k in ("exp_avg_sq_row", "exp_avg_res_row"):
                            st[k] = torch.zeros(grad.shape[:-1], device=grad.device)
                        for k in ("exp_avg_sq_col", "exp_avg_res_col"):
                            st[k] = torch.zeros(grad.shape[:-2] + grad.shape[-1:], device=grad.device)
                    else:
                        st["exp_avg_sq"] = torch.zeros_like(grad)
                st["step"] += 1
                update = grad ** 2 + group["eps"][0]
                if factored:
                    update = self._factored(st["exp_avg_sq_row"], st["exp_avg_sq_col"], update, b2, sd, nd).mul_(grad)
                else:
                    st["exp_avg_sq"].mul_(b2).add_(update, alpha=1.0 - b2)
                    update = st["exp_avg_sq"].rsqrt().mul_(grad)
                update.div_((self._rms(update, sd is not None) / group["clip_threshold"]).clamp_(min=1.0))
                st["exp_avg"].mul_(b1).add_(update, alpha=1 - b1)
                if factored:
                    res = (update - st["exp_avg"]) ** 2 + group["eps"][1]
                    update = self._factored(st["exp_avg_res_row"], st["exp_avg_res_col"], res, b3, sd, nd) \
                        .mul_(st["exp_avg"])
                else:
                    update = st["exp_avg"].clone()
                pf = p.float()
                if group["weight_decay"] != 0:
                    pf = pf - group["weight_decay"] * group["lr"] * pf
                p.copy_(pf - group["lr"] * update)
        return loss
