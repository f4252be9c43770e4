"""Capacity-based routers producing the `(combine weights, dispatch mask / slots)` the legacy layer consumes.
Parity: reference `colossalai/legacy/moe/layer/routers.py:1-470` (`MoeRouter`, `Top1Router`, `Top2Router`,
`TopKRouter`, `get_router_cls`)."""
from __future__ import annotations

import math
from abc import ABC
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...moe import moe_cumsum
from .manager import MOE_MANAGER


class MoeRouter(nn.Module, ABC):
    def __init__(self, k_value: int, capacity_factor_train: float, capacity_factor_eval: float, min_capacity: int,
                 noisy_func: Optional[Callable] = None, drop_tks: bool = True, use_kernel: bool = False) -> None:
        super().__init__()
        self.k_value = k_value
        self.capacity_factor_train, self.capacity_factor_eval = capacity_factor_train, capacity_factor_eval
        self.min_capacity, self.noisy_func, self.drop_tks = min_capacity, noisy_func, drop_tks
        self.use_kernel = use_kernel
        self._aux_loss = None
        self._z_loss = None

    def get_capacity(self, num_tokens: int, num_experts: int, ep_group=None) -> int:
        f = self.capacity_factor_train if self.training else self.capacity_factor_eval
        cap = math.floor(self.k_value * f * num_tokens / num_experts)
        cap += cap % 2
        cap = max(cap, self.min_capacity)
        assert cap > 0
        return int(cap)

    def set_aux_loss(self, router_probs: torch.Tensor, expert_indices: torch.Tensor, num_experts: int) -> None:
        """Switch / GShard load-balancing loss: E * sum_e f_e * P_e."""
        assert self._aux_loss is None
        if router_probs.dim() == expert_indices.dim() == 2:
            router_probs, expert_indices = router_probs.unsqueeze(0), expert_indices.unsqueeze(0)
        mask = F.one_hot(expert_indices, num_experts).max(dim=-2)[0]          # [g, s, e]
        tokens_per = mask.float().mean(dim=-2)
        prob_per = router_probs.float().mean(dim=-2)
        self._aux_loss = (tokens_per * prob_per).mean() * num_experts ** 2

    def set_z_loss(self, router_logits: torch.Tensor) -> None:
        assert self._z_loss is None
        self._z_loss = (torch.logsumexp(router_logits.float(), dim=-1) ** 2).mean()

    def pop_router_loss(self) -> None:
        MOE_MANAGER.add_loss(self._aux_loss if self._aux_loss is not None else 0.0,
                             self._z_loss if self._z_loss is not None else 0.0)
        self._aux_loss = self._z_loss = None


class TopKRouter(MoeRouter):
    """General top-k with per-expert capacity: returns `used_capacity`, combine weights `[s, e, c]` and the boolean
    dispatch mask `[s, e, c]`; the kernel path instead returns `(used_capacity, probs*mask [s, e], mask [s, e],
    slot [s, e], capacity)` for `MoeDispatch` / `MoeCombine`."""

    def __init__(self, k_value: int = 2, capacity_factor_train: float = 1.25, capacity_factor_eval: float = 2.0,
                 min_capacity: int = 4, noisy_func: Optional[Callable] = None, drop_tks: bool = True,
                 select_policy: str = "first", use_kernel: bool = False) -> None:
        super().__init__(k_value, capacity_factor_train, capacity_factor_eval, min_capacity, noisy_func, drop_tks,
                         use_kernel)
        assert select_policy in ("first", "random")
        self.select_policy = select_policy

    def forward(self, inputs: torch.Tensor, use_kernel: Optional[bool] = None, ep_group=None) -> Tuple:
        use_kernel = self.use_kernel if use_kernel is None else use_kernel
        if self.noisy_func is not None and self.training:
            inputs = self.noisy_func(inputs)
        assert inputs.dtype == torch.float, "router logits must be fp32"
        probs = F.softmax(inputs, dim=-1)
        s, e = probs.shape
        capacity = self.get_capacity(s, e, ep_group)
        topv, topi = probs.topk(self.k_value, dim=-1)
        self.set_aux_loss(probs, topi, e)
        self.set_z_loss(inputs)
        self.pop_router_loss()
        masks, slots = [], []
        used = torch.zeros(e, dtype=torch.long, device=probs.device)
        for j in range(self.k_value):                       # k-th choices queue behind all (k-1)-th choices
            m = F.one_hot(topi[:, j], e).to(torch.int32)
            if self.select_policy == "random" and self.training:
                order = torch.randperm(s, device=probs.device)
                inv = torch.empty_like(order)
                inv[order] = torch.arange(s, device=probs.device)
                rank = moe_cumsum(m[order].contiguous(), use_kernel=use_kernel)[inv]
            else:
                rank = moe_cumsum(m, use_kernel=use_kernel)
            rank = rank + used[None].to(rank.dtype)
            if self.drop_tks:
                m = m * (rank < capacity)
            used = used + m.sum(0)
            masks.append(m)
            slots.append(rank * m)
        mask = sum(masks)                                     # [s, e] 0/1 (an expert is chosen at most once per token)
        slot = sum(slots)
        weight = probs * mask
        if self.k_value > 1:                                  # renormalise the kept choices
            weight = weight / weight.sum(-1, keepdim=True).clamp(min=torch.finfo(weight.dtype).eps)
        used_capacity = mask.sum(0)
        if use_kernel:
            return used_capacity, weight, mask, slot.to(torch.int32), capacity
        sec = F.one_hot(slot.long().clamp(max=capacity - 1), capacity) * mask.unsqueeze(-1)
        combine = weight.unsqueeze(-1) * sec
        return used_capacity, combine, sec.bool()


class Top1Router(TopKRouter):
    def __init__(self, capacity_factor_train: float = 1.25, capacity_factor_eval: float = 2.0, min_capacity: int = 4,
                 select_policy: str = "first", noisy_func: Optional[Callable] = None, drop_tks: bool = True,
                 use_kernel: bool = False) -> None:
        super().__init__(1, capacity_factor_train, capacity_factor_eval, min_capacity, noisy_func, drop_tks,
                         select_policy, use_kernel)


class Top2Router(TopKRouter):
    def __init__(self, capacity_factor_train: float = 1.25, capacity_factor_eval: float = 2.0, min_capacity: int = 4,
                 noisy_func: Optional[Callable] = None, drop_tks: bool = True, use_kernel: bool = False) -> None:
        super().__init__(2, capacity_factor_train, capacity_factor_eval, min_capacity, noisy_func, drop_tks, "first",
                         use_kernel)


def get_router_cls(top_k: int, grouped: bool = False):
    if not grouped:
        return {1: Top1Router, 2: Top2Router}.get(top_k, TopKRouter)
    return TopKRouter
