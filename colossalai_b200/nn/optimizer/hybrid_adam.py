"""HybridAdam: CPU kernel for host-resident params, multi-tensor CUDA kernel for device-resident ones — the
default optimizer of the Gemini-style chunk manager.  Parity: reference `colossalai/nn/optimizer/hybrid_adam.py:60-191`."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ...ops import multi_tensor as mt
from ...ops._dispatch import use_native
from .cpu_adam import CPUAdam, cpu_adam_step
from .fused_adam import adam_reference_step

__all__ = ["HybridAdam"]


class HybridAdam(CPUAdam):
    num_fp32_shards_per_param = 2

    def __init__(self, model_params, lr: float = 1e-3, bias_correction: bool = True, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0, adamw_mode: bool = True,
                 nvme_offload_fraction: float = 0.0, nvme_offload_dir: Optional[str] = None, **defaults) -> None:
        super().__init__(model_params, lr, bias_correction, betas, eps, weight_decay, adamw_mode,
                         nvme_offload_fraction, nvme_offload_dir)
        self._tables: Dict[tuple, mt.TensorTable] = {}

    @torch.no_grad()
    def step(self, closure=None, div_scale: float = -1.0):
        loss = closure() if closure is not None else None
        inv_scale = 1.0 / div_scale if div_scale > 0 else 1.0
        self._pre_step("exp_avg", "exp_avg_sq")
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            gp, gg, gm, gv = [], [], [], []
            group_step = 0
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self._init_state(p, p.device)
                st["step"] += 1
                group_step = st["step"]
                if p.device.type == "cpu":
                    self._pre_update(p, "exp_avg", "exp_avg_sq")
                    cpu_adam_step(p.data, p.grad.data, st["exp_avg"], st["exp_avg_sq"], group["lr"], beta1, beta2,
                                  group["eps"], group["weight_decay"], st["step"], group["bias_correction"],
                                  self.adamw_mode, inv_scale)
                    self._post_update(p, "exp_avg", "exp_avg_sq")
                else:
                    gp.append(p.data)
                    gg.append(p.grad.data)
                    gm.append(st["exp_avg"])
                    gv.append(st["exp_avg_sq"])
            if gp:
                if use_native(gp[0]) and all(t.is_contiguous() for t in gp + gg):
                    key = mt.TensorTable.key_of(gp, gg, gm, gv)
                    tbl = self._tables.get(key)
                    if tbl is None:
                        if len(self._tables) > 16:
                            self._tables.clear()
                        tbl = self._tables[key] = mt.TensorTable(gp, gg, gm, gv, keepalive=True)
                    mt.adam(tbl, group["lr"], beta1, beta2, group["eps"], group["weight_decay"], group_step,
                            self.adamw_mode, group["bias_correction"], inv_scale)
                else:
                    adam_reference_step(gp, gg, gm, gv, group["lr"], beta1, beta2, group["eps"],
                                        group["weight_decay"], group_step, self.adamw_mode,
                                        group["bias_correction"], inv_scale)
        self._post_step()
        return loss
