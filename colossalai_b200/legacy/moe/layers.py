"""`SparseMLP`: gate -> capacity router -> dispatch (kernel or einsum) -> [all-to-all] -> experts -> [all-to-all] ->
combine.  Parity: reference `colossalai/legacy/moe/layer/layers.py:1-400` (uses `MoeDispatch` / `MoeCombine` at
`:185,205`, `AllToAll` for expert parallelism, optional load-balancer statistics)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...moe import AllToAll, MoeCombine, MoeDispatch
from .experts import MLPExperts
from .load_balance import LoadBalancer
from .manager import MOE_MANAGER
from .routers import get_router_cls
from .utils import get_noise_generator


class SparseMLP(nn.Module):
    def __init__(self, num_experts: int, hidden_size: int, intermediate_size: int, router_top_k: int = 1,
                 router_capacity_factor_train: float = 1.25, router_capacity_factor_eval: float = 2.0,
                 router_min_capacity: int = 4, router_noisy_policy: Optional[str] = None, router_drop_tks: bool = True,
                 mlp_activation: str = "gelu", mlp_gated: bool = False, enable_kernel: bool = False,
                 enable_load_balance: bool = False, load_balance_tolerance: float = 0.1,
                 load_balance_beam_width: int = 8, load_balance_group_swap_factor: float = 0.4) -> None:
        super().__init__()
        self.hidden_size, self.num_experts = hidden_size, num_experts
        self.enable_kernel = enable_kernel
        self.gate_weight = nn.Parameter(torch.empty(num_experts, hidden_size))
        nn.init.trunc_normal_(self.gate_weight, std=(0.1 / hidden_size) ** 0.5)
        self.router = get_router_cls(router_top_k)(
            capacity_factor_train=router_capacity_factor_train, capacity_factor_eval=router_capacity_factor_eval,
            min_capacity=router_min_capacity, noisy_func=get_noise_generator(router_noisy_policy, num_experts),
            drop_tks=router_drop_tks) if router_top_k in (1, 2) else get_router_cls(router_top_k)(
            k_value=router_top_k, capacity_factor_train=router_capacity_factor_train,
            capacity_factor_eval=router_capacity_factor_eval, min_capacity=router_min_capacity,
            noisy_func=get_noise_generator(router_noisy_policy, num_experts), drop_tks=router_drop_tks)
        self.experts = MLPExperts(num_experts, hidden_size, intermediate_size, activation=mlp_activation,
                                  gated=mlp_gated)
        self.ep_group = MOE_MANAGER.ep_group if MOE_MANAGER.parallel == "EP" else None
        self.ep_size = self.experts.ep_size
        self.num_local_experts = self.experts.num_local_experts
        self.enable_load_balance = enable_load_balance
        if enable_load_balance:
            self.load_balancer = LoadBalancer(self.experts, self.gate_weight, self.num_local_experts, num_experts,
                                              self.ep_group, tolerance=load_balance_tolerance,
                                              beam_width=load_balance_beam_width,
                                              group_swap_factor=load_balance_group_swap_factor)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        shape = inputs.shape
        tokens = inputs.reshape(-1, self.hidden_size)
        logits = F.linear(tokens.float(), self.gate_weight.float())
        if self.enable_load_balance and self.training:
            with torch.no_grad():
                self.load_balancer.update_load(F.one_hot(logits.argmax(-1), self.num_experts).sum(0))
        use_kernel = self.enable_kernel
        routed = self.router(logits, use_kernel=use_kernel, ep_group=self.ep_group)
        if use_kernel:
            _, weight, mask, slot, cap = routed
            x = MoeDispatch.apply(tokens, mask, slot, self.num_experts * cap)           # [e, c, h]
        else:
            _, combine, sec = routed
            cap = sec.shape[-1]
            x = torch.einsum("sec,sh->ech", sec.to(tokens.dtype), tokens)
        x = self._run_experts(x)                                                          # [e, c, h]
        if use_kernel:
            out = MoeCombine.apply(x.reshape(-1, self.hidden_size), weight.to(torch.float32), mask, slot,
                                   self.num_experts * cap)
        else:
            out = torch.einsum("sec,ech->sh", combine.to(x.dtype), x)
        return out.to(inputs.dtype).reshape(shape)

    def _run_experts(self, x: torch.Tensor) -> torch.Tensor:
        if self.ep_size == 1:
            return self.experts(x)
        e, c, h = x.shape
        # [ep, local, c, h] -> every rank receives the slots of ITS local experts from all ranks
        x = AllToAll.apply(x.reshape(self.ep_size, self.num_local_experts, c, h).contiguous(), self.ep_group)
        x = x.reshape(self.ep_size, self.num_local_experts, c, h).transpose(0, 1).reshape(self.num_local_experts, -1, h)
        x = self.experts(x)
        x = x.reshape(self.num_local_experts, self.ep_size, c, h).transpose(0, 1).contiguous()
        x = AllToAll.apply(x, self.ep_group)
        return x.reshape(e, c, h)
