"""Legacy `Trainer` with hooks.  Parity: reference `colossalai/legacy/trainer/_trainer.py` + `hooks/`."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from ..engine import Engine
from .hooks import BaseHook

__all__ = ["Trainer"]


class Trainer:
    def __init__(self, engine: Engine, timer=None, logger=None) -> None:
        self.engine = engine
        self.timer, self.logger = timer, logger
        self.cur_epoch = self.cur_step = 0
        self.max_epochs = self.max_steps = None
        self.steps_per_epoch = 0
        self.hooks: List[BaseHook] = []
        self.states = {}
        self._exit = False

    def _call_hooks(self, name: str, *args) -> None:
        for h in sorted(self.hooks, key=lambda h: h.priority):
            getattr(h, name)(self, *args)

    def _train_epoch(self, loader: Iterable) -> None:
        self.engine.train()
        self._call_hooks("before_train_epoch")
        it = iter(loader)
        for _ in range(self.steps_per_epoch):
            if self._exit:
                break
            self._call_hooks("before_train_iter")
            self.engine.zero_grad()
            out, label, loss = self.engine.execute_schedule(it)
            self.engine.step()
            self._call_hooks("after_train_iter", out, label, loss)
            self.cur_step += 1
            if self.max_steps is not None and self.cur_step >= self.max_steps:
                self._exit = True
        self._call_hooks("after_train_epoch")
        self.cur_epoch += 1

    @torch.no_grad()
    def _eval(self, loader: Iterable) -> None:
        self.engine.eval()
        self._call_hooks("before_test_epoch")
        it = iter(loader)
        for _ in range(len(loader)):
            self._call_hooks("before_test_iter")
            out, label, loss = self.engine.execute_schedule(it, forward_only=True)
            self._call_hooks("after_test_iter", out, label, loss)
        self._call_hooks("after_test_epoch")

    def fit(self, train_dataloader, epochs: int, max_steps: Optional[int] = None, test_dataloader=None,
            test_interval: int = 1, hooks: Optional[List[BaseHook]] = None, display_progress: bool = False) -> None:
        self.max_epochs, self.max_steps = epochs, max_steps
        self.steps_per_epoch = len(train_dataloader)
        self.hooks = list(hooks or [])
        self._call_hooks("before_train")
        for epoch in range(self.cur_epoch, epochs):
            if self._exit:
                break
            self._train_epoch(train_dataloader)
            if test_dataloader is not None and (epoch + 1) % test_interval == 0:
                self._eval(test_dataloader)
        self._call_hooks("after_train")

    def evaluate(self, test_dataloader, hooks: Optional[List[BaseHook]] = None) -> None:
        self.hooks = list(hooks or self.hooks)
        self._call_hooks("before_test")
        self._eval(test_dataloader)
        self._call_hooks("after_test")
