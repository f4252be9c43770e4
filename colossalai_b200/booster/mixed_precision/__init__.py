"""Mixed-precision strategies for plugins that do not control precision themselves.
Parity: reference `colossalai/booster/mixed_precision/*.py` (`FP16TorchMixedPrecision`, `FP16ApexMixedPrecision`,
`FP16NaiveMixedPrecision`, `BF16MixedPrecision`, `FP8MixedPrecision`, `mixed_precision_factory`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor
from torch.optim import Optimizer

from ...interface import ModelWrapper, OptimizerWrapper

__all__ = ["MixedPrecision", "FP16TorchMixedPrecision", "FP16ApexMixedPrecision", "FP16NaiveMixedPrecision",
           "BF16MixedPrecision", "FP8MixedPrecision", "mixed_precision_factory", "TorchAMPOptimizer", "TorchAMPModule"]


class MixedPrecision(ABC):
    @abstractmethod
    def configure(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None
                  ) -> Tuple[nn.Module, OptimizerWrapper, Callable]:
        ...


class TorchAMPOptimizer(OptimizerWrapper):
    def __init__(self, optim: Optimizer, init_scale: float = 2.0**16, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5, growth_interval: int = 2000, device: str = "cuda") -> None:
        super().__init__(optim)
        self.scaler = torch.amp.GradScaler(device, init_scale=init_scale, growth_factor=growth_factor,
                                           backoff_factor=backoff_factor, growth_interval=growth_interval)

    def backward(self, loss: Tensor, *args, **kwargs) -> None:
        self.scaler.scale(loss).backward(*args, **kwargs)

    def step(self, *args, **kwargs):
        out = self.scaler.step(self.optim, *args, **kwargs)
        self.scaler.update()
        return out

    def scale_loss(self, loss: Tensor) -> Tensor:
        return self.scaler.scale(loss)

    def unscale_grad(self) -> None:
        self.scaler.unscale_(self.optim)

    def clip_grad_by_value(self, clip_value: float, *args, **kwargs) -> None:
        self.unscale_grad()
        super().clip_grad_by_value(clip_value, *args, **kwargs)

    def clip_grad_by_norm(self, max_norm, norm_type=2.0, error_if_nonfinite=False, *args, **kwargs):
        self.unscale_grad()
        return super().clip_grad_by_norm(max_norm, norm_type, error_if_nonfinite, *args, **kwargs)


class TorchAMPModule(ModelWrapper):
    def __init__(self, module: nn.Module, dtype: torch.dtype = torch.float16) -> None:
        super().__init__(module)
        self.dtype = dtype

    def forward(self, *args, **kwargs):
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        with torch.autocast(dev, dtype=self.dtype if dev == "cuda" else torch.bfloat16):
            return self.module(*args, **kwargs)


class FP16TorchMixedPrecision(MixedPrecision):
    def __init__(self, init_scale: float = 2.0**16, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000) -> None:
        self.kw = dict(init_scale=init_scale, growth_factor=growth_factor, backoff_factor=backoff_factor,
                       growth_interval=growth_interval)

    def configure(self, model, optimizer=None, criterion=None):
        model = TorchAMPModule(model)
        if optimizer is not None:
            optimizer = TorchAMPOptimizer(optimizer, device="cuda" if torch.cuda.is_available() else "cpu", **self.kw)
        if criterion is not None:
            inner = criterion
            criterion = lambda *a, **k: inner(*a, **k)  # noqa: E731
        return model, optimizer, criterion


class BF16MixedPrecision(MixedPrecision):
    def configure(self, model, optimizer=None, criterion=None):
        model = TorchAMPModule(model, dtype=torch.bfloat16)
        if optimizer is not None:
            optimizer = OptimizerWrapper(optimizer)
        return model, optimizer, criterion


class FP16NaiveMixedPrecision(MixedPrecision):
    """Cast the model to fp16 and wrap the optimizer with fp32 master weights + dynamic loss scale."""

    def __init__(self, log_num_zeros_in_grad: bool = False, initial_scale: float = 2**16, growth_factor: float = 2,
                 backoff_factor: float = 0.5, growth_interval: int = 1000, hysteresis: int = 2,
                 max_scale: float = 2**32, verbose: bool = False, max_norm: float = 0.0) -> None:
        self.kw = dict(initial_scale=initial_scale, growth_factor=growth_factor, backoff_factor=backoff_factor,
                       growth_interval=growth_interval, hysteresis=hysteresis, max_scale=max_scale, max_norm=max_norm)

    def configure(self, model, optimizer=None, criterion=None):
        from ...amp import MixedPrecisionOptimizer

        model = model.half()
        if optimizer is not None:
            optimizer = MixedPrecisionOptimizer(optimizer, model, precision="fp16", **self.kw)
        return model, optimizer, criterion


class FP16ApexMixedPrecision(FP16NaiveMixedPrecision):
    """apex is not a dependency here; the `fp16_apex` name maps to the naive fp16 strategy (same semantics:
    fp16 model + fp32 master weights + dynamic loss scaling)."""

    def __init__(self, opt_level: str = "O1", **kw) -> None:
        super().__init__(**{k: v for k, v in kw.items() if k in ("initial_scale", "max_norm")})
        self.opt_level = opt_level


class FP8MixedPrecision(MixedPrecision):
    """bf16 storage + fp8 (e4m3 fwd / e5m2 bwd) GEMMs in every nn.Linear via `quantization.fp8.linear_fp8`."""

    def configure(self, model, optimizer=None, criterion=None):
        from ...quantization.fp8_hook import convert_linear_to_fp8

        model = convert_linear_to_fp8(model.to(torch.bfloat16))
        if optimizer is not None:
            from ...amp import MixedPrecisionOptimizer

            optimizer = MixedPrecisionOptimizer(optimizer, model, precision="bf16")
        return model, optimizer, criterion


_mixed_precision_mapping = {
    "fp16": FP16TorchMixedPrecision,
    "fp16_apex": FP16ApexMixedPrecision,
    "fp16_naive": FP16NaiveMixedPrecision,
    "bf16": BF16MixedPrecision,
    "fp8": FP8MixedPrecision,
}


def mixed_precision_factory(mixed_precision_type: str) -> MixedPrecision:
    if mixed_precision_type in _mixed_precision_mapping:
        return _mixed_precision_mapping[mixed_precision_type]()
    raise ValueError(f"mixed precision {mixed_precision_type} is not supported; "
                     f"choose from {list(_mixed_precision_mapping.keys())}")
