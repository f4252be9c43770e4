"""Functional entry points to ZeRO 1/2/3 without a Booster.  Parity: reference `colossalai/zero/wrapper.py:10,52`."""
from __future__ import annotations

from copy import copy
from typing import Dict, Optional

import torch
import torch.nn as nn

from .gemini import GeminiDDP
from .gemini.gemini_optimizer import GeminiOptimizer
from .low_level import LowLevelZeroOptimizer

__all__ = ["zero_model_wrapper", "zero_optim_wrapper"]


def zero_model_wrapper(model: nn.Module, zero_stage: int = 1, gemini_config: Optional[Dict] = None,
                       verbose: bool = False) -> nn.Module:
    assert zero_stage in (1, 2, 3), "The stage of ZeRO should be 1, 2 or 3"
    if gemini_config is None:
        gemini_config = {}
    if zero_stage in (1, 2):
        wrapped = model
    else:
        wrapped = GeminiDDP(model, **gemini_config, verbose=verbose)
    setattr(wrapped, "_colo_zero_stage", zero_stage)
    return wrapped


def zero_optim_wrapper(model: nn.Module, optimizer: torch.optim.Optimizer, initial_scale: float = 2**16,
                       growth_factor: float = 2, backoff_factor: float = 0.5, growth_interval: int = 1000,
                       hysteresis: int = 2, min_scale: float = 1, max_scale: float = 2**32, max_norm: float = 0.0,
                       norm_type: float = 2.0, optim_config: Optional[Dict] = None, verbose: bool = False):
    assert hasattr(model, "_colo_zero_stage"), "You should use `zero_model_wrapper` first"
    zero_stage = getattr(model, "_colo_zero_stage")
    assert norm_type == 2.0, "Current ZeRO optimizers only support 'norm_type=2'"
    config = copy(optim_config) if optim_config is not None else {}
    config.update(initial_scale=initial_scale, growth_factor=growth_factor, backoff_factor=backoff_factor,
                  growth_interval=growth_interval, hysteresis=hysteresis, min_scale=min_scale, max_scale=max_scale)
    if zero_stage in (1, 2):
        config["partition_grad"] = zero_stage == 2
        config["clip_grad_norm"] = max_norm
        return LowLevelZeroOptimizer(optimizer, **config, verbose=verbose)
    config["max_norm"] = max_norm
    return GeminiOptimizer(optimizer, model, **config, verbose=verbose)
