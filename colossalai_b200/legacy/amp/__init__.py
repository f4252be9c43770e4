"""Legacy mixed-precision entry point.  Parity: reference `colossalai/legacy/amp/{__init__.py:1-60, amp_type.py}`
(`AMP_TYPE.{TORCH, APEX, NAIVE}`, `convert_to_amp(model, optimizer, criterion, mode, amp_config)`).  APEX O2-style and
NAIVE both map onto our fp32-master `MixedPrecisionOptimizer`; TORCH is autocast + `GradScaler`."""
from __future__ import annotations

from enum import Enum
from typing import Optional

import torch
import torch.nn as nn

__all__ = ["AMP_TYPE", "convert_to_amp"]


class AMP_TYPE(Enum):
    APEX = "apex"
    TORCH = "torch"
    NAIVE = "naive"


class _AutocastModel(nn.Module):
    def __init__(self, model: nn.Module, dtype: torch.dtype) -> None:
        super().__init__()
        self.model, self.dtype = model, dtype

    def forward(self, *a, **k):
        dev = next(self.model.parameters()).device.type
        with torch.autocast(device_type=dev, dtype=self.dtype):
            return self.model(*a, **k)


def convert_to_amp(model: nn.Module, optimizer, criterion=None, mode: AMP_TYPE = AMP_TYPE.NAIVE,
                   amp_config: Optional[dict] = None):
    cfg = dict(amp_config or {})
    if mode == AMP_TYPE.TORCH:
        from ...booster.mixed_precision import FP16TorchMixedPrecision

        mp = FP16TorchMixedPrecision(**cfg)
        model, optimizer, criterion = mp.configure(model, optimizer, criterion)
        return model, optimizer, criterion
    from ...amp.naive_amp.mixed_precision_optimizer import MixedPrecisionOptimizer

    dtype = cfg.pop("dtype", torch.float16)
    model = model.to(dtype)
    precision = "bf16" if dtype == torch.bfloat16 else "fp16"
    optimizer = MixedPrecisionOptimizer(optimizer, model, precision=precision, **cfg)
    return model, optimizer, criterion
