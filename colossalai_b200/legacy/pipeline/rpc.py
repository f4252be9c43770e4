"""RPC-driven pipeline engines: a master process owns the schedule, every stage lives in a `PipelineWorker` held
through an `RRef`; micro-batch activations travel stage to stage as RPC arguments, each worker keeps the autograd
graph of its micro-batches and back-propagates when it receives the output gradient.

`FillDrainPipelineEngine` runs all forwards then all backwards (GPipe); `OneFOneBPipelineEngine` bounds the number of
in-flight micro-batches per stage to `num_stages - stage` like the non-interleaved 1F1B schedule.

Parity: reference `colossalai/legacy/pipeline/rpc/{_pipeline_base.py (WorkerBase, PipelineEngineBase),
_pipeline_schedule.py (FillDrainWorker/Engine, OneFOneBWorker/Engine), utils.py}`.
"""
from __future__ import annotations

import threading
from typing import Any, Callable, Dict, List, Optional

import torch
import torch.distributed.rpc as rpc
import torch.nn as nn

__all__ = ["PipelineWorker", "FillDrainPipelineEngine", "OneFOneBPipelineEngine", "rpc_is_initialized"]


def rpc_is_initialized() -> bool:
    try:
        rpc.get_worker_info()
        return True
    except RuntimeError:
        return False


class PipelineWorker:
    """Lives on the stage's process.  Thread-safe: RPC calls of different micro-batches may interleave."""

    def __init__(self, partition_fn: Callable, stage: int, num_stages: int, device: str = "cpu",
                 criterion: Optional[Callable] = None, checkpoint: bool = False) -> None:
        self.stage, self.num_stages = stage, num_stages
        self.device = torch.device(device)
        self.module: nn.Module = partition_fn(stage).to(self.device)
        self.criterion, self.checkpoint = criterion, checkpoint
        self.lock = threading.Lock()
        self.saved: Dict[int, Any] = {}
        self.optimizer = None

    # ---- forward of one micro-batch; returns detached output (or the loss on the last stage when labels are given)
    def forward(self, mb: int, x, labels=None, forward_only: bool = False):
        x = x.to(self.device) if torch.is_tensor(x) else x
        with self.lock:
            if forward_only:
                with torch.no_grad():
                    out = self.module(x)
                if labels is not None and self.criterion is not None:
                    out = self.criterion(out, labels.to(self.device))
                return out.detach().cpu()
            inp = x.detach().requires_grad_(x.is_floating_point()) if torch.is_tensor(x) else x
            if self.checkpoint:
                from torch.utils.checkpoint import checkpoint as ck

                out = ck(self.module, inp, use_reentrant=False)
            else:
                out = self.module(inp)
            if labels is not None and self.criterion is not None:
                out = self.criterion(out, labels.to(self.device))
            self.saved[mb] = (inp, out)
            return out.detach().cpu()

    # ---- backward of one micro-batch; returns the gradient w.r.t. the stage input (None on stage 0)
    def backward(self, mb: int, grad_out=None):
        with self.lock:
            inp, out = self.saved.pop(mb)
            if grad_out is None:
                out.backward()
            else:
                out.backward(grad_out.to(self.device))
            g = inp.grad if (torch.is_tensor(inp) and inp.requires_grad) else None
            return None if g is None else g.detach().cpu()

    def in_flight(self) -> int:
        return len(self.saved)

    def init_optimizer(self, optimizer_class: type, **kwargs) -> None:
        self.optimizer = optimizer_class(self.module.parameters(), **kwargs)

    def step(self, scale: float = 1.0) -> None:
        if scale != 1.0:
            for p in self.module.parameters():
                if p.grad is not None:
                    p.grad.mul_(scale)
        self.optimizer.step()
        self.optimizer.zero_grad()

    def state_dict(self):
        return {k: v.cpu() for k, v in self.module.state_dict().items()}

    def grads(self):
        return {n: (None if p.grad is None else p.grad.detach().cpu()) for n, p in self.module.named_parameters()}


class _EngineBase:
    schedule = "fill_drain"

    def __init__(self, partition_fn: Callable, stage_num: int, num_microbatches: int, device: str = "cpu",
                 chunk: int = 1, criterion: Optional[Callable] = None, metric: Optional[Callable] = None,
                 checkpoint: bool = False, data_process_func: Optional[Callable] = None,
                 worker_names: Optional[List[str]] = None) -> None:
        assert chunk == 1, "interleaved chunks are provided by the current pipeline schedules (pipeline/schedule)"
        assert rpc_is_initialized(), "call torch.distributed.rpc.init_rpc first"
        self.stage_num, self.num_microbatches = stage_num, num_microbatches
        self.criterion, self.metric, self.data_process_func = criterion, metric, data_process_func
        names = worker_names or [f"work{i}" for i in range(stage_num)]
        self.workers = [rpc.remote(names[s], PipelineWorker,
                                   args=(partition_fn, s, stage_num, device, criterion if s == stage_num - 1 else None,
                                         checkpoint)) for s in range(stage_num)]

    def initialize_optimizer(self, optimizer_class: type, **kwargs) -> None:
        for w in self.workers:
            w.rpc_sync().init_optimizer(optimizer_class, **kwargs)

    def step(self) -> None:
        for f in [w.rpc_async().step(1.0 / self.num_microbatches) for w in self.workers]:
            f.wait()

    def remote_parameters_grads(self) -> List[dict]:
        return [w.rpc_sync().grads() for w in self.workers]

    def _fwd_chain(self, mb: int, x, label, forward_only: bool):
        out = x
        for s, w in enumerate(self.workers):
            last = s == self.stage_num - 1
            out = w.rpc_sync().forward(mb, out, label if last else None, forward_only)
        return out

    def _bwd_chain(self, mb: int) -> None:
        g = None
        for w in reversed(self.workers):
            g = w.rpc_sync().backward(mb, g)

    def forward_backward(self, batch: torch.Tensor, labels: Optional[torch.Tensor] = None, forward_only: bool = False):
        if self.data_process_func is not None:
            batch = self.data_process_func(batch)
        mbs = list(batch.chunk(self.num_microbatches))
        lbs = list(labels.chunk(self.num_microbatches)) if labels is not None else [None] * len(mbs)
        outs: List[Any] = [None] * len(mbs)
        threads: List[threading.Thread] = []
        sem = threading.Semaphore(self._max_in_flight())

        def run(i: int) -> None:
            with sem:
                outs[i] = self._fwd_chain(i, mbs[i], lbs[i], forward_only)
                if not forward_only and self.schedule == "1f1b":
                    self._bwd_chain(i)

        for i in range(len(mbs)):       # micro-batches enter the pipe in order; stages overlap through the RPC threads
            t = threading.Thread(target=run, args=(i,))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
        if not forward_only and self.schedule == "fill_drain":
            bts = [threading.Thread(target=self._bwd_chain, args=(i,)) for i in range(len(mbs))]
            for t in bts:
                t.start()
            for t in bts:
                t.join()
        return outs

    def _max_in_flight(self) -> int:
        return self.num_microbatches


class FillDrainPipelineEngine(_EngineBase):
    schedule = "fill_drain"


class OneFOneBPipelineEngine(_EngineBase):
    schedule = "1f1b"

    def _max_in_flight(self) -> int:
        return min(self.num_microbatches, self.stage_num)
