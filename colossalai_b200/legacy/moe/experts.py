"""Batched expert MLPs, expert-parallel over `MOE_MANAGER.ep_group`.
Parity: reference `colossalai/legacy/moe/layer/experts.py:1-160` (`MLPExperts`: `wi` / `wo` (+ gate) parameter stacks
of the LOCAL experts, `ep_group` tagging via the moe-tensor API)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...tensor.moe_tensor import set_moe_tensor_ep_group
from .manager import MOE_MANAGER


class MLPExperts(nn.Module):
    def __init__(self, num_experts: int, hidden_size: int, intermediate_size: int, activation: str = "gelu",
                 gated: bool = False, drop_rate: float = 0.0, use_kernel: bool = False) -> None:
        super().__init__()
        ep = MOE_MANAGER.ep_size if MOE_MANAGER.parallel == "EP" else 1
        assert num_experts % ep == 0, f"{num_experts} experts not divisible by ep size {ep}"
        self.num_total_experts, self.num_local_experts, self.ep_size = num_experts, num_experts // ep, ep
        self.gated, self.act_name, self.drop_rate = gated, activation, drop_rate
        n = self.num_local_experts
        self.wi = nn.Parameter(torch.empty(n, hidden_size, intermediate_size * (2 if gated else 1)))
        self.wo = nn.Parameter(torch.empty(n, intermediate_size, hidden_size))
        self.reset_parameters()
        if ep > 1:
            for p in self.parameters():
                set_moe_tensor_ep_group(p, MOE_MANAGER.ep_group)

    @torch.no_grad()
    def reset_parameters(self) -> None:
        # every rank seeds by its expert-parallel rank so different experts get different weights
        nn.init.trunc_normal_(self.wi, std=math.sqrt(0.1 / self.wi.shape[1]))
        nn.init.trunc_normal_(self.wo, std=math.sqrt(0.1 / self.wo.shape[1]))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """`x` [local_experts, tokens, hidden] -> same shape."""
        h = torch.bmm(x, self.wi.to(x.dtype))
        if self.gated:
            h = ops.glu(h, self.act_name)
        else:
            h = ops.get_activation(self.act_name)(h)
        if self.drop_rate > 0 and self.training:
            h = nn.functional.dropout(h, self.drop_rate)
        return torch.bmm(h, self.wo.to(x.dtype))
