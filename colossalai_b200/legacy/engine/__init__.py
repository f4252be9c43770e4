"""Legacy `Engine`: model + optimizer + criterion behind train()/eval()/__call__/backward/step.
Parity: reference `colossalai/legacy/engine/_base_engine.py` (+ gradient handlers collapsed into the dp all-reduce)."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Optimizer

__all__ = ["Engine", "DataParallelGradientHandler"]


class DataParallelGradientHandler:
    """Averages gradients over a data-parallel group after backward (bucketed all-reduce)."""

    def __init__(self, model: nn.Module, group=None) -> None:
        self.model, self.group = model, group

    def handle_gradient(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=self.group)
        flat.div_(dist.get_world_size(self.group))
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


class Engine:
    def __init__(self, model: nn.Module, optimizer: Optimizer, criterion: Optional[Callable] = None,
                 gradient_handlers: Optional[List] = None, clip_grad_norm: float = 0.0, verbose: bool = True) -> None:
        self._model, self._optimizer, self._criterion = model, optimizer, criterion
        self._gradient_handlers = gradient_handlers or []
        self._clip_grad_norm = clip_grad_norm
        self.training = True
        self.verbose = verbose

    @property
    def model(self) -> nn.Module:
        return self._model

    @property
    def optimizer(self) -> Optimizer:
        return self._optimizer

    @property
    def criterion(self):
        return self._criterion

    def train(self) -> None:
        self.training = True
        self._model.train()

    def eval(self) -> None:
        self.training = False
        self._model.eval()

    def zero_grad(self) -> None:
        self._optimizer.zero_grad()

    def __call__(self, *args, **kwargs):
        return self._model(*args, **kwargs)

    def backward(self, loss: torch.Tensor) -> None:
        loss.backward()

    def step(self):
        for h in self._gradient_handlers:
            h.handle_gradient()
        if self._clip_grad_norm > 0:
            torch.nn.utils.clip_grad_norm_(self._model.parameters(), self._clip_grad_norm)
        return self._optimizer.step()

    def execute_schedule(self, data_iter: Iterable, forward_only: bool = False, return_loss: bool = True):
        """One non-pipelined step: forward (+criterion) (+backward); returns (output, label, loss)."""
        batch = next(data_iter)
        data, label = (batch["data"], batch.get("label")) if isinstance(batch, dict) else batch
        out = self(data)
        loss = self._criterion(out, label) if (self._criterion is not None and return_loss) else None
        if not forward_only and loss is not None:
            self.backward(loss)
        return out, label, loss
