"""FusedSGD (multi-tensor kernel).  Parity: reference `colossalai/nn/optimizer/fused_sgd.py:77`."""
from __future__ import annotations

from typing import Dict

import torch
from torch.optim.optimizer import Optimizer, required

from ...ops import multi_tensor as mt
from ...ops._dispatch import use_native

__all__ = ["FusedSGD"]


class FusedSGD(Optimizer):
    def __init__(self, params, lr=required, momentum: float = 0, dampening: float = 0, weight_decay: float = 0,
                 nesterov: bool = False, wd_after_momentum: bool = False) -> None:
        if lr is not required and lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))
        self.wd_after_momentum = wd_after_momentum
        self._tables: Dict[tuple, mt.TensorTable] = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, gs, ms = [], [], []
            first_run = False
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "momentum_buffer" not in st:
                    first_run = True
                    st["momentum_buffer"] = torch.zeros_like(p, dtype=torch.float32)
                ps.append(p.data)
                gs.append(p.grad.data)
                ms.append(st["momentum_buffer"])
            if not ps:
                continue
            if use_native(ps[0]) and all(t.is_contiguous() for t in ps + gs):
                mt.sgd(mt.TensorTable(ps, gs, ms), group["lr"], group["momentum"], group["dampening"],
                       group["weight_decay"], group["nesterov"], first_run, self.wd_after_momentum)
            else:
                for p, g, buf in zip(ps, gs, ms):
                    d = g.float()
                    pf = p.float()
                    if group["weight_decay"] != 0 and not self.wd_after_momentum:
                        d = d + group["weight_decay"] * pf
                    if group["momentum"] != 0:
                        if first_run:
                            buf.copy_(d)
                        else:
                            buf.mul_(group["momentum"]).add_(d, alpha=1 - group["dampening"])
                        d = d + group["momentum"] * buf if group["nesterov"] else buf
                    if group["weight_decay"] != 0 and self.wd_after_momentum:
                        d = d + group["weight_decay"] * pf
                    p.copy_(pf - group["lr"] * d)
        return loss
