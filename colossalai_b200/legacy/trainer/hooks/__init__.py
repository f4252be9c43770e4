"""Trainer hooks.  Parity: reference `colossalai/legacy/trainer/hooks/{_base,_log,_metric,_checkpoint,_lr_scheduler}_hook.py`."""
from __future__ import annotations

import os
from typing import Optional

import torch

__all__ = ["BaseHook", "LossHook", "LogMetricByEpochHook", "LRSchedulerHook", "SaveCheckpointHook", "ThroughputHook"]


class BaseHook:
    priority = 5

    def __init__(self, priority: Optional[int] = None) -> None:
        if priority is not None:
            self.priority = priority

    def before_train(self, trainer): ...
    def after_train(self, trainer): ...
    def before_train_epoch(self, trainer): ...
    def after_train_epoch(self, trainer): ...
    def before_train_iter(self, trainer): ...
    def after_train_iter(self, trainer, output, label, loss): ...
    def before_test(self, trainer): ...
    def after_test(self, trainer): ...
    def before_test_epoch(self, trainer): ...
    def after_test_epoch(self, trainer): ...
    def before_test_iter(self, trainer): ...
    def after_test_iter(self, trainer, output, label, loss): ...


class LossHook(BaseHook):
    """Running mean of the training / test loss in `trainer.states['metrics']`."""

    priority = 0

    def before_train_epoch(self, trainer):
        trainer.states.setdefault("metrics", {})["train_loss"] = [0.0, 0]

    def after_train_iter(self, trainer, output, label, loss):
        if loss is not None:
            m = trainer.states["metrics"]["train_loss"]
            m[0] += float(loss)
            m[1] += 1

    def before_test_epoch(self, trainer):
        trainer.states.setdefault("metrics", {})["test_loss"] = [0.0, 0]

    def after_test_iter(self, trainer, output, label, loss):
        if loss is not None:
            m = trainer.states["metrics"]["test_loss"]
            m[0] += float(loss)
            m[1] += 1


class LogMetricByEpochHook(BaseHook):
    priority = 10

    def __init__(self, logger=None, interval: int = 1, priority: int = 10) -> None:
        super().__init__(priority)
        self.logger, self.interval = logger, interval
        self.history = []

    def _emit(self, trainer, mode: str) -> None:
        vals = {k: (v[0] / max(v[1], 1)) for k, v in trainer.states.get("metrics", {}).items() if k.startswith(mode)}
        msg = f"[Epoch {trainer.cur_epoch} / {mode}] " + " | ".join(f"{k} = {v:.5f}" for k, v in vals.items())
        self.history.append((trainer.cur_epoch, mode, vals))
        if self.logger is not None:
            self.logger.info(msg, ranks=[0])

    def after_train_epoch(self, trainer):
        if trainer.cur_epoch % self.interval == 0:
            self._emit(trainer, "train")

    def after_test_epoch(self, trainer):
        self._emit(trainer, "test")


class LRSchedulerHook(BaseHook):
    priority = 1

    def __init__(self, lr_scheduler, by_epoch: bool = True, priority: int = 1) -> None:
        super().__init__(priority)
        self.lr_scheduler, self.by_epoch = lr_scheduler, by_epoch

    def after_train_iter(self, trainer, output, label, loss):
        if not self.by_epoch:
            self.lr_scheduler.step()

    def after_train_epoch(self, trainer):
        if self.by_epoch:
            self.lr_scheduler.step()


class SaveCheckpointHook(BaseHook):
    priority = 10

    def __init__(self, interval: int = 1, checkpoint_dir: str = "./ckpt", model=None, save_by_iter: bool = False,
                 priority: int = 10) -> None:
        super().__init__(priority)
        self.interval, self.dir, self.model, self.by_iter = interval, checkpoint_dir, model, save_by_iter

    def _save(self, trainer, tag: str) -> None:
        import torch.distributed as dist

        if dist.is_initialized() and dist.get_rank() != 0:
            return
        os.makedirs(self.dir, exist_ok=True)
        model = self.model if self.model is not None else trainer.engine.model
        torch.save({"model": model.state_dict(), "optimizer": trainer.engine.optimizer.state_dict(),
                    "epoch": trainer.cur_epoch, "step": trainer.cur_step}, os.path.join(self.dir, f"{tag}.pt"))

    def after_train_iter(self, trainer, output, label, loss):
        if self.by_iter and (trainer.cur_step + 1) % self.interval == 0:
            self._save(trainer, f"iter_{trainer.cur_step + 1}")

    def after_train_epoch(self, trainer):
        if not self.by_iter and (trainer.cur_epoch + 1) % self.interval == 0:
            self._save(trainer, f"epoch_{trainer.cur_epoch + 1}")


class ThroughputHook(BaseHook):
    priority = 10

    def __init__(self, tokens_per_step: int = 0, priority: int = 10) -> None:
        super().__init__(priority)
        self.tokens_per_step = tokens_per_step
        self._t0 = None
        self.samples = []

    def before_train_iter(self, trainer):
        import time

        self._t0 = time.perf_counter()

    def after_train_iter(self, trainer, output, label, loss):
        import time

        self.samples.append(self.tokens_per_step / max(time.perf_counter() - self._t0, 1e-9))
