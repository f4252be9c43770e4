"""Precision mixins.  Parity: reference `colossalai/amp/naive_amp/mixed_precision_mixin/{base,bf16,fp16}.py`."""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from ..grad_scaler import DynamicGradScaler

__all__ = ["MixedPrecisionMixin", "BF16MixedPrecisionMixin", "FP16MixedPrecisionMixin"]


class MixedPrecisionMixin(ABC):
    dtype: torch.dtype

    @abstractmethod
    def pre_backward(self, loss: Tensor, *args, **kwargs) -> Tensor:
        ...

    @abstractmethod
    def pre_backward_by_grad(self, tensor: Tensor, grad: Tensor):
        ...

    @abstractmethod
    def should_skip_step(self) -> bool:
        ...

    @abstractmethod
    def pre_zero_grad(self) -> None:
        ...

    @abstractmethod
    def get_grad_div_scale(self) -> float:
        ...


class BF16MixedPrecisionMixin(MixedPrecisionMixin):
    dtype = torch.bfloat16

    def pre_backward(self, loss: Tensor, *a, **k) -> Tensor:
        return loss

    def pre_backward_by_grad(self, tensor: Tensor, grad: Tensor):
        return grad

    def should_skip_step(self) -> bool:
        return False

    def pre_zero_grad(self) -> None:
        pass

    def get_grad_div_scale(self) -> float:
        return 1.0


class FP16MixedPrecisionMixin(MixedPrecisionMixin):
    dtype = torch.float16

    class OptimState(Enum):
        SCALED = 0
        UNSCALED = 1

    def __init__(self, initial_scale: float = 2**16, min_scale: float = 1, growth_factor: float = 2,
                 backoff_factor: float = 0.5, growth_interval: int = 1000, hysteresis: int = 2,
                 max_scale: float = 2**32) -> None:
        self.grad_scaler = DynamicGradScaler(initial_scale, growth_factor, backoff_factor, growth_interval, min_scale,
                                             max_scale, hysteresis)
        self.optim_state = self.OptimState.UNSCALED
        self.found_overflow = torch.zeros(1, dtype=torch.float, device=self.grad_scaler.scale.device)

    @property
    def loss_scale(self) -> float:
        return self.grad_scaler.scale.item()

    @abstractmethod
    def check_local_overflow(self) -> bool:
        ...

    def check_overflow(self) -> bool:
        self.found_overflow.fill_(1.0 if self.check_local_overflow() else 0.0)
        if dist.is_initialized():
            dist.all_reduce(self.found_overflow, op=dist.ReduceOp.MAX)
        return self.found_overflow.item() > 0

    def pre_backward(self, loss: Tensor, *a, **k) -> Tensor:
        loss = self.loss_scale * loss
        self.optim_state = self.OptimState.SCALED
        return loss

    def pre_backward_by_grad(self, tensor: Tensor, grad: Tensor):
        self.optim_state = self.OptimState.SCALED
        return grad

    def should_skip_step(self) -> bool:
        found_inf = self.check_overflow()
        self.grad_scaler.update(found_inf)
        if found_inf:
            self.optim_state = self.OptimState.UNSCALED
        return found_inf

    def pre_zero_grad(self) -> None:
        pass

    def get_grad_div_scale(self) -> float:
        assert self.optim_state == self.OptimState.SCALED, "grads should be scaled before clipping"
        self.optim_state = self.OptimState.UNSCALED
        return self.loss_scale
