"""Trainer skeletons.  Parity: reference `coati/trainer/base.py:1-218` (`SLTrainer` for supervised-style objectives,
`OLTrainer` for online RL: collect -> update loops) and `trainer/utils.py` (`all_reduce_mean`, `CycledDataLoader`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from colossalai_b200.booster import Booster


def all_reduce_mean(t: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = t.detach().clone()
        dist.all_reduce(t)
        t /= dist.get_world_size()
    return t


def is_rank_0() -> bool:
    return not dist.is_initialized() or dist.get_rank() == 0


class CycledDataLoader:
    """Endless `next()` over a finite dataloader (online trainers draw prompts by count, not by epoch)."""

    def __init__(self, dataloader: Iterable) -> None:
        self.dataloader, self._it, self.count = dataloader, None, 0

    def next(self):
        if self._it is None:
            self._it = iter(self.dataloader)
        self.count += 1
        try:
            return next(self._it)
        except StopIteration:
            self._it = iter(self.dataloader)
            return next(self._it)


class _TrainerBase(ABC):
    def __init__(self, booster: Optional[Booster], optimizer, lr_scheduler=None, accumulation_steps: int = 1,
                 device=None, callbacks: Optional[List] = None) -> None:
        from .callbacks import CallbackList

        self.callbacks = CallbackList(callbacks)
        self.booster, self.optimizer, self.lr_scheduler = booster, optimizer, lr_scheduler
        self.accumulation_steps = max(1, accumulation_steps)
        self.device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                                 else torch.device("cpu"))
        self.history: List[Dict[str, float]] = []
        self._micro = 0

    def _backward_and_maybe_step(self, loss: torch.Tensor) -> bool:
        loss = loss / self.accumulation_steps
        if self.booster is not None:
            self.booster.backward(loss, self.optimizer)
        else:
            loss.backward()
        self._micro += 1
        if self._micro % self.accumulation_steps == 0:
            self.optimizer.step()
            self.optimizer.zero_grad()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            return True
        return False

    def add_callbacks(self, *callbacks) -> "_TrainerBase":
        """Attach callbacks after construction (every concrete trainer gets them without a signature change)."""
        self.callbacks.callbacks.extend(callbacks)
        return self

    def _to_device(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        return {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def log(self, **metrics: float) -> None:
        self.history.append({k: float(v) for k, v in metrics.items()})


class SLTrainer(_TrainerBase):
    """Epoch loop over a dataloader calling `_train_step(batch) -> (loss, metrics)`."""

    def __init__(self, booster, max_epochs: int, model: nn.Module, optimizer, lr_scheduler=None,
                 accumulation_steps: int = 1, device=None, callbacks: Optional[List] = None) -> None:
        super().__init__(booster, optimizer, lr_scheduler, accumulation_steps, device, callbacks)
        self.max_epochs, self.model = max_epochs, model

    @abstractmethod
    def _train_step(self, batch: Dict[str, torch.Tensor]):
        ...

    def _eval_step(self, batch: Dict[str, torch.Tensor]) -> Dict[str, float]:
        with torch.no_grad():
            loss, metrics = self._train_step(batch)
        return {"loss": float(loss), **metrics}

    def fit(self, train_dataloader: Iterable, eval_dataloader: Optional[Iterable] = None) -> List[Dict[str, float]]:
        self.callbacks.on_fit_start(self)
        for epoch in range(self.max_epochs):
            self.model.train()
            self.callbacks.on_epoch_start(self, epoch)
            for batch in train_dataloader:
                batch = self._to_device(batch)
                self.callbacks.on_batch_start(self, batch)
                loss, metrics = self._train_step(batch)
                self._backward_and_maybe_step(loss)
                self.log(epoch=epoch, loss=all_reduce_mean(loss.detach()), **metrics)
                self.callbacks.on_batch_end(self, batch, self.history[-1])
            if eval_dataloader is not None:
                self.model.eval()
                agg: Dict[str, float] = {}
                n = 0
                for batch in eval_dataloader:
                    for k, v in self._eval_step(self._to_device(batch)).items():
                        agg[k] = agg.get(k, 0.0) + v
                    n += 1
                self.log(epoch=epoch, **{f"eval_{k}": v / max(n, 1) for k, v in agg.items()})
            self.callbacks.on_epoch_end(self, epoch)
        self.callbacks.on_fit_end(self)
        return self.history


class OLTrainer(_TrainerBase):
    """`fit(prompts, num_episodes, num_collect_steps, num_update_steps)`: collect rollouts, then update on them."""

    def __init__(self, booster, optimizer, lr_scheduler=None, accumulation_steps: int = 1, device=None,
                 callbacks: Optional[List] = None) -> None:
        super().__init__(booster, optimizer, lr_scheduler, accumulation_steps, device, callbacks)

    @abstractmethod
    def _collect(self, prompts: Dict[str, torch.Tensor]) -> None:
        ...

    @abstractmethod
    def _update(self) -> Dict[str, float]:
        ...

    def _after_episode(self) -> None:
        pass

    def fit(self, prompt_dataloader: Iterable, num_episodes: int = 1, num_collect_steps: int = 1,
            num_update_steps: int = 1) -> List[Dict[str, float]]:
        prompts = CycledDataLoader(prompt_dataloader)
        self.callbacks.on_fit_start(self)
        for episode in range(num_episodes):
            self.callbacks.on_epoch_start(self, episode)
            for _ in range(num_collect_steps):
                batch = self._to_device(prompts.next())
                self.callbacks.on_collect_start(self)
                self._collect(batch)
                self.callbacks.on_collect_end(self, batch)
            for _ in range(num_update_steps):
                self.callbacks.on_update_start(self)
                self.log(episode=episode, **self._update())
                self.callbacks.on_update_end(self, self.history[-1])
            self._after_episode()
            self.callbacks.on_epoch_end(self, episode)
        self.callbacks.on_fit_end(self)
        return self.history
