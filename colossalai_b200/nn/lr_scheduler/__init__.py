"""LR schedulers.  Parity: reference `colossalai/nn/lr_scheduler/{cosine,linear,multistep,onecycle,poly,torch,
delayed}.py` (cosine / linear / multistep / one-cycle / polynomial + warm-up and delayed wrappers)."""
from __future__ import annotations

import math
from typing import List, Optional

from torch.optim.lr_scheduler import CosineAnnealingLR as _CosineAnnealingLR
from torch.optim.lr_scheduler import ExponentialLR as _ExponentialLR
from torch.optim.lr_scheduler import LambdaLR as _LambdaLR
from torch.optim.lr_scheduler import MultiplicativeLR as _MultiplicativeLR
from torch.optim.lr_scheduler import MultiStepLR as _MultiStepLR
from torch.optim.lr_scheduler import OneCycleLR as _OneCycleLR
from torch.optim.lr_scheduler import StepLR as _StepLR
from torch.optim.lr_scheduler import _LRScheduler

__all__ = ["CosineAnnealingLR", "CosineAnnealingWarmupLR", "FlatAnnealingLR", "FlatAnnealingWarmupLR", "LinearWarmupLR",
           "MultiStepLR", "MultiStepWarmupLR", "OneCycleLR", "PolynomialLR", "PolynomialWarmupLR", "LambdaLR",
           "MultiplicativeLR", "StepLR", "ExponentialLR", "DelayerScheduler", "WarmupScheduler",
           "WarmupDelayerScheduler"]


class _enable_get_lr_call:
    def __init__(self, o):
        self.o = o

    def __enter__(self):
        self.o._get_lr_called_within_step = True
        return self

    def __exit__(self, *a):
        self.o._get_lr_called_within_step = False


class DelayerScheduler(_LRScheduler):
    """Keep the initial lr for `delay_epochs`, then hand over to `after_scheduler`."""

    def __init__(self, optimizer, delay_epochs: int, after_scheduler: _LRScheduler, last_epoch: int = -1):
        if delay_epochs < 0:
            raise ValueError(f"delay_epochs must >= 0, got {delay_epochs}")
        self.delay_epochs = delay_epochs
        self.after_scheduler = after_scheduler
        self.finished = False
        super().__init__(optimizer, last_epoch)

    def state_dict(self):
        sd = {k: v for k, v in self.__dict__.items() if k not in ("optimizer", "after_scheduler")}
        sd["after_scheduler_dict"] = self.after_scheduler.state_dict()
        return sd

    def load_state_dict(self, sd):
        self.after_scheduler.load_state_dict(sd.pop("after_scheduler_dict"))
        self.__dict__.update(sd)

    def get_lr(self):
        if self.last_epoch >= self.delay_epochs:
            if not self.finished:
                self.after_scheduler.base_lrs = self.base_lrs
                self.finished = True
            with _enable_get_lr_call(self.after_scheduler):
                return self.after_scheduler.get_lr()
        return self.base_lrs

    def step(self, epoch=None):
        if self.finished:
            if epoch is None:
                self.after_scheduler.step(None)
                self._last_lr = self.after_scheduler.get_last_lr()
            else:
                self.after_scheduler.step(epoch - self.delay_epochs)
                self._last_lr = self.after_scheduler.get_last_lr()
        else:
            return super().step(epoch)


class WarmupScheduler(_LRScheduler):
    """Linear warm-up for `warmup_epochs` steps, then `after_scheduler`."""

    def __init__(self, optimizer, warmup_epochs: int, after_scheduler: _LRScheduler, last_epoch: int = -1):
        self.warmup_epochs = int(warmup_epochs)
        self.after_scheduler = after_scheduler
        self.finished = False
        super().__init__(optimizer, last_epoch)

    def state_dict(self):
        sd = {k: v for k, v in self.__dict__.items() if k not in ("optimizer", "after_scheduler")}
        sd["after_scheduler_dict"] = self.after_scheduler.state_dict()
        return sd

    def load_state_dict(self, sd):
        self.after_scheduler.load_state_dict(sd.pop("after_scheduler_dict"))
        self.__dict__.update(sd)

    def get_lr(self):
        if self.last_epoch >= self.warmup_epochs:
            if not self.finished:
                self.after_scheduler.base_lrs = self.base_lrs
                self.finished = True
            return self.after_scheduler.get_lr()
        return [(self.last_epoch + 1) / self.warmup_epochs * lr for lr in self.base_lrs]

    def step(self, epoch=None):
        if self.finished:
            if epoch is None:
                self.after_scheduler.step(None)
                self._last_lr = self.after_scheduler.get_last_lr()
            else:
                self.after_scheduler.step(epoch - self.warmup_epochs)
                self._last_lr = self.after_scheduler.get_last_lr()
        else:
            return super().step(epoch)


class WarmupDelayerScheduler(_LRScheduler):
    """Warm-up, then hold, then `after_scheduler`."""

    def __init__(self, optimizer, warmup_epochs: int, delay_epochs: int, after_scheduler, last_epoch: int = -1):
        if delay_epochs < 0 or warmup_epochs < 0:
            raise ValueError("warmup_epochs / delay_epochs must be >= 0")
        self.warmup_epochs, self.delay_epochs = warmup_epochs, delay_epochs
        self.after_scheduler = after_scheduler
        self.finished = False
        super().__init__(optimizer, last_epoch)

    def state_dict(self):
        sd = {k: v for k, v in self.__dict__.items() if k not in ("optimizer", "after_scheduler")}
        sd["after_scheduler_dict"] = self.after_scheduler.state_dict()
        return sd

    def load_state_dict(self, sd):
        self.after_scheduler.load_state_dict(sd.pop("after_scheduler_dict"))
        self.__dict__.update(sd)

    def get_lr(self):
        if self.last_epoch >= self.warmup_epochs + self.delay_epochs:
            if not self.finished:
                self.after_scheduler.base_lrs = self.base_lrs
                self.finished = True
            with _enable_get_lr_call(self.after_scheduler):
                return self.after_scheduler.get_lr()
        if self.last_epoch >= self.warmup_epochs:
            return self.base_lrs
        return [(self.last_epoch + 1) / self.warmup_epochs * lr for lr in self.base_lrs]

    def step(self, epoch=None):
        if self.finished:
            if epoch is None:
                self.after_scheduler.step(None)
                self._last_lr = self.after_scheduler.get_last_lr()
            else:
                self.after_scheduler.step(epoch - self.warmup_epochs)
                self._last_lr = self.after_scheduler.get_last_lr()
        else:
            return super().step(epoch)


class CosineAnnealingLR(_CosineAnnealingLR):
    def __init__(self, optimizer, total_steps: int, eta_min: float = 0, last_epoch: int = -1, **kwargs):
        super().__init__(optimizer, total_steps, eta_min=eta_min, last_epoch=last_epoch)


class CosineAnnealingWarmupLR(WarmupScheduler):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int = 0, eta_min: float = 0.0, last_epoch: int = -1):
        base = _CosineAnnealingLR(optimizer, total_steps - warmup_steps, eta_min=eta_min, last_epoch=last_epoch)
        super().__init__(optimizer, warmup_steps, base, last_epoch=last_epoch)


class FlatAnnealingLR(DelayerScheduler):
    def __init__(self, optimizer, total_steps: int, pct_start: float = 0.72, last_epoch: int = -1, **kwargs):
        if not (0.0 <= pct_start <= 1.0):
            raise ValueError(f"pct_start must >= 0.0 and <= 1.0, got {pct_start}")
        flat = int(total_steps * pct_start)
        base = _CosineAnnealingLR(optimizer, total_steps - flat)
        super().__init__(optimizer, flat, base, last_epoch=last_epoch)


class FlatAnnealingWarmupLR(WarmupDelayerScheduler):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int = 0, pct_start: float = 0.72,
                 eta_min: float = 0, last_epoch: int = -1, **kwargs):
        if not (0.0 <= pct_start <= 1.0):
            raise ValueError(f"pct_start must >= 0.0 and <= 1.0, got {pct_start}")
        flat = int((total_steps - warmup_steps) * pct_start)
        base = _CosineAnnealingLR(optimizer, total_steps - warmup_steps - flat, eta_min=eta_min)
        super().__init__(optimizer, warmup_steps, flat, base, last_epoch=last_epoch)


class LinearWarmupLR(_LRScheduler):
    """Linear warm-up then linear decay to zero."""

    def __init__(self, optimizer, total_steps: int, warmup_steps: int = 0, last_epoch: int = -1, **kwargs):
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        super().__init__(optimizer, last_epoch=last_epoch)

    def get_lr(self):
        if self.last_epoch < self.warmup_steps:
            return [(self.last_epoch + 1) / (self.warmup_steps + 1) * lr for lr in self.base_lrs]
        return [(self.total_steps - self.last_epoch) / max(self.total_steps - self.warmup_steps, 1) * lr
                for lr in self.base_lrs]


class MultiStepLR(_MultiStepLR):
    def __init__(self, optimizer, total_steps: int = None, milestones: List[int] = None, gamma: float = 0.1,
                 last_epoch: int = -1, **kwargs):
        super().__init__(optimizer, milestones, gamma=gamma, last_epoch=last_epoch)


class MultiStepWarmupLR(WarmupScheduler):
    def __init__(self, optimizer, total_steps: int = None, warmup_steps: int = 0, milestones: List[int] = None,
                 gamma: float = 0.1, last_epoch: int = -1, **kwargs):
        if len(milestones) == 0:
            raise ValueError("milestones cannot be empty")
        ms = [v - warmup_steps for v in milestones if v >= warmup_steps]
        base = _MultiStepLR(optimizer, ms, gamma=gamma)
        super().__init__(optimizer, warmup_steps, base, last_epoch=last_epoch)


class OneCycleLR(_OneCycleLR):
    def __init__(self, optimizer, total_steps: int, pct_start=0.3, anneal_strategy="cos", cycle_momentum=True,
                 base_momentum=0.85, max_momentum=0.95, div_factor=25.0, final_div_factor=10000.0, last_epoch=-1,
                 **kwargs):
        max_lrs = [g["lr"] for g in optimizer.param_groups]
        super().__init__(optimizer, max_lrs, total_steps=total_steps, pct_start=pct_start,
                         anneal_strategy=anneal_strategy, cycle_momentum=cycle_momentum, base_momentum=base_momentum,
                         max_momentum=max_momentum, div_factor=div_factor, final_div_factor=final_div_factor,
                         last_epoch=last_epoch)


class PolynomialLR(_LRScheduler):
    def __init__(self, optimizer, total_steps: int, end_lr: float = 0.0001, power: float = 1.0, last_epoch: int = -1,
                 **kwargs):
        if end_lr < 0:
            raise ValueError(f"end_lr must >= 0, got {end_lr}")
        self.total_steps, self.end_lr, self.power = total_steps, end_lr, power
        super().__init__(optimizer, last_epoch=last_epoch)

    def get_lr(self):
        return self._get_closed_form_lr()

    def _get_closed_form_lr(self):
        return [(b - self.end_lr) * ((1 - min(self.last_epoch, self.total_steps) / self.total_steps) ** self.power)
                + self.end_lr for b in self.base_lrs]


class PolynomialWarmupLR(WarmupScheduler):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int = 0, end_lr: float = 0.0001, power: float = 1.0,
                 last_epoch: int = -1, **kwargs):
        base = PolynomialLR(optimizer, total_steps - warmup_steps, end_lr=end_lr, power=power)
        super().__init__(optimizer, warmup_steps, base, last_epoch=last_epoch)


class LambdaLR(_LambdaLR):
    def __init__(self, optimizer, total_steps=None, lr_lambda=None, last_epoch: int = -1):
        super().__init__(optimizer, lr_lambda, last_epoch=last_epoch)


class MultiplicativeLR(_MultiplicativeLR):
    def __init__(self, optimizer, total_steps=None, lr_lambda=None, last_epoch: int = -1):
        super().__init__(optimizer, lr_lambda, last_epoch=last_epoch)


class StepLR(_StepLR):
    def __init__(self, optimizer, total_steps=None, step_size: int = 1, gamma: float = 0.1, last_epoch: int = -1):
        super().__init__(optimizer, step_size, gamma=gamma, last_epoch=last_epoch)


class ExponentialLR(_ExponentialLR):
    def __init__(self, optimizer, total_steps=None, gamma: float = 1.0, last_epoch: int = -1):
        super().__init__(optimizer, gamma, last_epoch=last_epoch)
