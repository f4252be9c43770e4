"""Loss scalers.  Parity: reference `colossalai/amp/naive_amp/grad_scaler/{base,constant,dynamic}_grad_scaler.py`."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, Optional

import torch
from torch import Tensor

__all__ = ["BaseGradScaler", "ConstantGradScaler", "DynamicGradScaler"]


class BaseGradScaler(ABC):
    def __init__(self, initial_scale: float, verbose: bool = False) -> None:
        assert initial_scale > 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self._scale = torch.tensor([initial_scale], dtype=torch.float32, device=dev)
        self._verbose = verbose

    @property
    def scale(self) -> Tensor:
        return self._scale

    @property
    def inv_scale(self) -> Tensor:
        return self._scale.double().reciprocal().float()

    def state_dict(self) -> Dict:
        return {"scale": self.scale}

    def load_state_dict(self, state_dict: Dict) -> None:
        self._scale = state_dict["scale"]

    @abstractmethod
    def update(self, overflow: bool) -> None:
        ...


class ConstantGradScaler(BaseGradScaler):
    def update(self, overflow: bool) -> None:
        pass


class DynamicGradScaler(BaseGradScaler):
    def __init__(self, initial_scale: float = 2**16, growth_factor: float = 2, backoff_factor: float = 0.5,
                 growth_interval: int = 1000, min_scale: Optional[float] = None, max_scale: Optional[float] = None,
                 hysteresis: int = 2, verbose: bool = False) -> None:
        super().__init__(initial_scale, verbose)
        self._min_scale = None if min_scale is None else torch.tensor([min_scale], device=self._scale.device)
        self._max_scale = None if max_scale is None else torch.tensor([max_scale], device=self._scale.device)
        self._growth_factor, self._backoff_factor = growth_factor, backoff_factor
        self._growth_interval, self._hysteresis = growth_interval, hysteresis
        self._growth_step, self._hysteresis_step = 0, 0
        assert growth_factor > 1 and 0 < backoff_factor < 1 and hysteresis >= 0
        if min_scale is not None and max_scale is not None:
            assert min_scale <= max_scale

    def update(self, overflow: bool) -> None:
        if overflow:
            self._hysteresis_step += 1
            self._growth_step = 0
            if self._hysteresis_step >= self._hysteresis:
                self._backoff_scale()
        else:
            self._growth_step += 1
            if self._growth_step == self._growth_interval:
                self._growth_step = 0
                self._hysteresis_step = 0
                self._grow_scale()

    def _backoff_scale(self) -> None:
        self._scale = self._scale * self._backoff_factor
        if self._min_scale is not None:
            self._scale = torch.max(self._scale, self._min_scale)

    def _grow_scale(self) -> None:
        self._scale = self._scale * self._growth_factor
        if self._max_scale is not None:
            self._scale = torch.min(self._scale, self._max_scale)

    def state_dict(self) -> Dict:
        return dict(scale=self._scale, growth_factor=self._growth_factor, backoff_factor=self._backoff_factor,
                    hysteresis=self._hysteresis)

    def load_state_dict(self, sd: Dict) -> None:
        self._scale = sd["scale"].to(self._scale.device)
        self._growth_factor, self._backoff_factor = sd["growth_factor"], sd["backoff_factor"]
        self._hysteresis = sd["hysteresis"]
