"""DistributedLamb: LAMB whose trust-ratio norms are reduced over the TP group (sharded params) and the ZeRO dp group.
Parity: reference `colossalai/nn/optimizer/distributed_lamb.py`."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from ...interface.optimizer import DistributedOptim
from ...tensor.d_tensor import is_distributed_tensor

__all__ = ["DistributedLamb"]


class DistributedLamb(DistributedOptim):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, bias_correction=True) -> None:
        assert lr >= 0 and eps >= 0
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      bias_correction=bias_correction))
        self.shard_to_working_param: Dict = {}
        self.tp_size = self.dp_size = 1
        self.tp_group = self.dp_group = None
        self.is_zero = False

    def _post_cast(self) -> None:
        for a, v in (("shard_to_working_param", {}), ("tp_size", 1), ("dp_size", 1), ("tp_group", None),
                     ("dp_group", None), ("is_zero", False)):
            if not hasattr(self, a):
                setattr(self, a, v)
        for g in self.param_groups:
            g.setdefault("bias_correction", True)

    def setup_distributed(self, tp_group=None, dp_group=None, shard_to_working_param: Optional[Dict] = {},
                          padding_map=None, is_zero: Optional[bool] = False) -> None:
        self.tp_group, self.dp_group = tp_group, dp_group
        self.tp_size = dist.get_world_size(tp_group) if tp_group is not None else 1
        self.dp_size = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.shard_to_working_param = shard_to_working_param or {}
        self.is_zero = bool(is_zero)

    def _working(self, p):
        return self.shard_to_working_param.get(id(p), p)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.float()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
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
                sq = torch.stack([pf.pow(2).sum(), update.pow(2).sum()])
                wp = self._working(p)
                if self.tp_size > 1 and is_distributed_tensor(wp):
                    dist.all_reduce(sq, group=self.tp_group)
                if self.is_zero and self.dp_size > 1:
                    dist.all_reduce(sq, group=self.dp_group)
                w_norm, u_norm = sq[0].sqrt(), sq[1].sqrt()
                trust = torch.where((w_norm > 0) & (u_norm > 0), w_norm / u_norm, torch.ones_like(w_norm))
                p.copy_(pf - group["lr"] * trust * update)
        return loss
