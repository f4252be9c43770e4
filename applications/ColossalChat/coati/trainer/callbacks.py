"""Trainer callbacks (reference `coati/trainer/callbacks/{base,performance_evaluator,save_checkpoint}.py`).

A callback sees the trainer at the natural boundaries of both loop families (`SLTrainer`: epoch / batch;
`OLTrainer`: episode / collect / update) and is free to keep state.  `PerformanceEvaluator` turns those boundaries into
throughput numbers: samples/s and tokens/s of the whole job, TFLOP/s per device from the usual transformer estimate
(6 x parameters x tokens for a training step, 2 x for generation, + the attention term), device-timed when a CUDA
stream is active and all-reduced over the job so that every rank reports the same (slowest-rank) figure."""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist

__all__ = ["Callback", "CallbackList", "PerformanceEvaluator", "SaveCheckpoint", "MetricsLogger"]


class Callback:
    def on_fit_start(self, trainer) -> None: ...
    def on_fit_end(self, trainer) -> None: ...
    def on_epoch_start(self, trainer, epoch: int) -> None: ...
    def on_epoch_end(self, trainer, epoch: int) -> None: ...
    def on_batch_start(self, trainer, batch: Dict[str, Any]) -> None: ...
    def on_batch_end(self, trainer, batch: Dict[str, Any], metrics: Dict[str, float]) -> None: ...
    def on_collect_start(self, trainer) -> None: ...
    def on_collect_end(self, trainer, batch: Dict[str, Any]) -> None: ...
    def on_update_start(self, trainer) -> None: ...
    def on_update_end(self, trainer, metrics: Dict[str, float]) -> None: ...


class CallbackList(Callback):
    def __init__(self, callbacks: Optional[List[Callback]] = None) -> None:
        self.callbacks = list(callbacks or [])

    def __getattribute__(self, name: str):
        if name.startswith("on_"):
            cbs = object.__getattribute__(self, "callbacks")

            def fan_out(*args, **kwargs):
                for cb in cbs:
                    getattr(cb, name)(*args, **kwargs)
            return fan_out
        return object.__getattribute__(self, name)


def _sync_time() -> float:
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.perf_counter()


def _max_over_ranks(x: float) -> float:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


class PerformanceEvaluator(Callback):
    """`PerformanceEvaluator(num_params, num_layers=L, hidden_size=H, ignore_steps=1)`; read `.summary()` after `fit`."""

    def __init__(self, num_params: int, num_layers: int = 0, hidden_size: int = 0, ignore_steps: int = 0,
                 enable_grad_checkpoint: bool = False, token_key: str = "input_ids", mask_key: str = "attention_mask") -> None:
        self.num_params, self.num_layers, self.hidden_size = num_params, num_layers, hidden_size
        self.ignore_steps, self.token_key, self.mask_key = ignore_steps, token_key, mask_key
        self.fwd_bwd_factor = 4 if enable_grad_checkpoint else 3     # forward = 1, backward = 2, recompute = +1
        self.reset()

    def reset(self) -> None:
        self.steps = 0
        self.train_time = self.train_tokens = self.train_samples = self.train_flops = 0.0
        self.gen_time = self.gen_tokens = self.gen_flops = 0.0
        self._t0: Optional[float] = None

    # ---- flops model
    def _flops(self, tokens: float, seq_len: float, factor: float) -> float:
        dense = 2.0 * self.num_params * tokens
        attn = 4.0 * self.num_layers * self.hidden_size * seq_len * tokens if self.num_layers else 0.0   # QK^T + PV
        return factor * (dense + attn)

    def _count(self, batch: Dict[str, Any]):
        ids = batch.get(self.token_key)
        if ids is None:
            ids = next(v for v in batch.values() if torch.is_tensor(v) and v.dim() >= 2)
        mask = batch.get(self.mask_key)
        tokens = float(mask.sum()) if torch.is_tensor(mask) else float(ids.numel())
        return float(ids.shape[0]), tokens, float(ids.shape[-1])

    # ---- supervised loops
    def on_batch_start(self, trainer, batch) -> None:
        self._t0 = _sync_time()

    def on_batch_end(self, trainer, batch, metrics) -> None:
        dt = _max_over_ranks(_sync_time() - self._t0)
        self.steps += 1
        if self.steps <= self.ignore_steps:
            return
        n, tok, seq = self._count(batch)
        world = dist.get_world_size() if dist.is_initialized() else 1
        self.train_time += dt
        self.train_samples += n * world
        self.train_tokens += tok * world
        self.train_flops += self._flops(tok, seq, self.fwd_bwd_factor)

    # ---- online loops
    def on_collect_start(self, trainer) -> None:
        self._t0 = _sync_time()

    def on_collect_end(self, trainer, batch) -> None:
        dt = _max_over_ranks(_sync_time() - self._t0)
        n, tok, seq = self._count(batch)
        world = dist.get_world_size() if dist.is_initialized() else 1
        self.gen_time += dt
        self.gen_tokens += tok * world
        self.gen_flops += self._flops(tok, seq / 2, 1.0)         # causal decoding: half the context on average

    def on_update_start(self, trainer) -> None:
        self._t0 = _sync_time()

    def on_update_end(self, trainer, metrics) -> None:
        self.train_time += _max_over_ranks(_sync_time() - self._t0)
        self.steps += 1

    def summary(self) -> Dict[str, float]:
        out = {"steps": float(self.steps)}
        if self.train_time > 0:
            out.update(train_samples_per_s=self.train_samples / self.train_time,
                       train_tokens_per_s=self.train_tokens / self.train_time,
                       train_tflops_per_device=self.train_flops / self.train_time / 1e12)
        if self.gen_time > 0:
            out.update(generate_tokens_per_s=self.gen_tokens / self.gen_time,
                       generate_tflops_per_device=self.gen_flops / self.gen_time / 1e12)
        return out

    def on_fit_end(self, trainer) -> None:
        trainer.performance = self.summary()


class SaveCheckpoint(Callback):
    """Every `interval` epochs / episodes: `booster.save_model` (sharded) + optimizer + a small `progress.json`."""

    def __init__(self, path: str, interval: int = 1, model_attr: str = "model", save_optimizer: bool = True) -> None:
        self.path, self.interval, self.model_attr, self.save_optimizer = Path(path), max(1, interval), model_attr, save_optimizer
        self.saved: List[str] = []

    def _save(self, trainer, tag: str) -> None:
        model = getattr(trainer, self.model_attr, None) or getattr(trainer, "actor", None)
        target = self.path / tag
        if trainer.booster is not None:
            trainer.booster.save_model(model, str(target / "model"), shard=True)
            if self.save_optimizer:
                trainer.booster.save_optimizer(trainer.optimizer, str(target / "optimizer"), shard=True)
        else:
            target.mkdir(parents=True, exist_ok=True)
            torch.save(model.state_dict(), target / "model.pt")
            if self.save_optimizer:
                torch.save(trainer.optimizer.state_dict(), target / "optimizer.pt")
        if not dist.is_initialized() or dist.get_rank() == 0:
            (target / "progress.json").write_text(json.dumps({"tag": tag, "history_len": len(trainer.history)}))
        self.saved.append(tag)

    def on_epoch_end(self, trainer, epoch: int) -> None:
        if (epoch + 1) % self.interval == 0:
            self._save(trainer, f"epoch_{epoch}")


class MetricsLogger(Callback):
    """Appends every logged step to a jsonl file on rank 0 (what the reference sends to wandb / tensorboard)."""

    def __init__(self, path: str) -> None:
        self.path = Path(path)
        self._seen = 0

    def _flush(self, trainer) -> None:
        if dist.is_initialized() and dist.get_rank() != 0:
            return
        self.path.parent.mkdir(parents=True, exist_ok=True)
        with self.path.open("a") as f:
            for rec in trainer.history[self._seen:]:
                f.write(json.dumps(rec) + "\n")
        self._seen = len(trainer.history)

    def on_batch_end(self, trainer, batch, metrics) -> None:
        self._flush(trainer)

    def on_update_end(self, trainer, metrics) -> None:
        self._flush(trainer)

    def on_fit_end(self, trainer) -> None:
        self._flush(trainer)
