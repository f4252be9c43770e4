"""CPUAdam: host-memory Adam/AdamW driven by the native C++ kernel (AVX-512 / AVX2 dispatched at run time).
Parity: reference `colossalai/nn/optimizer/cpu_adam.py` + NVMe base (`nvme_optimizer.py`)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ...kernel import loader
from ...ops._dtypes import code
from .fused_adam import adam_reference_step
from .nvme_optimizer import NVMeOptimizer

__all__ = ["CPUAdam", "cpu_adam_step"]

_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        lib = loader.load("cb200_cpu_adam")
        lib.cb_cpu_sumsq.restype = ctypes.c_double
        lib.cb_cpu_isa.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def cpu_adam_step(p: torch.Tensor, g: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float,
                  beta1: float, beta2: float, eps: float, weight_decay: float, step: int, bias_correction: bool,
                  adamw: bool, inv_scale: float = 1.0, lp: Optional[torch.Tensor] = None) -> None:
    """One Adam update of a host tensor (any of fp32/fp16/bf16 param & grad; fp32 moments)."""
    assert p.device.type == "cpu" and g.device.type == "cpu"
    assert exp_avg.dtype == torch.float32 and exp_avg_sq.dtype == torch.float32
    if not (p.is_contiguous() and g.is_contiguous()):
        raise ValueError("cpu_adam_step needs contiguous tensors")
    lib = _get_lib()
    f = ctypes.c_float
    rc = lib.cb_cpu_adam_step(loader.ptr(p), code(p.dtype), loader.ptr(g), code(g.dtype), loader.ptr(exp_avg),
                              loader.ptr(exp_avg_sq), loader.ptr(lp), code(lp.dtype) if lp is not None else 0,
                              ctypes.c_int64(p.numel()), f(lr), f(beta1), f(beta2), f(eps), f(weight_decay),
                              int(step), int(bias_correction), int(adamw), f(inv_scale))
    if rc != 0:
        raise RuntimeError(f"cb_cpu_adam_step failed with code {rc}")


class CPUAdam(NVMeOptimizer):
    """Adam for parameters living in host memory (optionally with states on NVMe)."""

    num_fp32_shards_per_param = 3   # fp32 param + exp_avg + exp_avg_sq (NVMe accounting)

    def __init__(self, model_params, lr: float = 1e-3, bias_correction: bool = True, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0, adamw_mode: bool = True,
                 nvme_offload_fraction: float = 0.0, nvme_offload_dir: Optional[str] = None) -> None:
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, bias_correction=bias_correction)
        super().__init__(model_params, defaults, nvme_offload_fraction, nvme_offload_dir)
        self.adamw_mode = adamw_mode

    def _init_state(self, p: torch.Tensor, device: torch.device):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32, device=device)
            st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32, device=device)
            self._post_state_init(p)
        return st

    @torch.no_grad()
    def step(self, closure=None, div_scale: float = -1.0):
        loss = closure() if closure is not None else None
        inv_scale = 1.0 / div_scale if div_scale > 0 else 1.0
        self._pre_step("exp_avg", "exp_avg_sq")
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self._init_state(p, p.device)
                st["step"] += 1
                self._pre_update(p, "exp_avg", "exp_avg_sq")
                if p.device.type == "cpu":
                    cpu_adam_step(p.data, p.grad.data, st["exp_avg"], st["exp_avg_sq"], group["lr"], beta1, beta2,
                                  group["eps"], group["weight_decay"], st["step"], group["bias_correction"],
                                  self.adamw_mode, inv_scale)
                else:
                    adam_reference_step([p.data], [p.grad.data], [st["exp_avg"]], [st["exp_avg_sq"]], group["lr"],
                                        beta1, beta2, group["eps"], group["weight_decay"], st["step"],
                                        self.adamw_mode, group["bias_correction"], inv_scale)
                self._post_update(p, "exp_avg", "exp_avg_sq")
        self._post_step()
        return loss
