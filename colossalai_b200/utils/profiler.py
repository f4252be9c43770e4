"""Profiling helpers: profile contexts, a device-timed performance evaluator and a collective-traffic profiler.

Parity: reference `examples/language/performance_evaluator.py:39-178` (`get_profile_context`, `PerformanceEvaluator`) and
`legacy/utils/profiler/legacy/comm_profiler.py:56-318` (`CommProfiler`).  Differences by design: step time is measured
with CUDA events on the training stream and reduced with MAX over ranks (the reference uses un-synchronised host time
and the mean)."""
from __future__ import annotations

import time
from contextlib import contextmanager, nullcontext
from typing import Dict, Optional

import torch
import torch.distributed as dist

__all__ = ["get_profile_context", "PerformanceEvaluator", "CommProfiler"]


class _NsysGate:
    """cudaProfilerStart/Stop at the warm-up / active step boundaries (run under `nsys --capture-range=cudaProfilerApi`)."""

    def __init__(self, warmup_steps: int, active_steps: int) -> None:
        self.warmup, self.active, self.n = warmup_steps, active_steps, 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def step(self) -> None:
        if torch.cuda.is_available():
            if self.n == self.warmup:
                torch.cuda.cudart().cudaProfilerStart()
            elif self.n == self.warmup + self.active:
                torch.cuda.cudart().cudaProfilerStop()
        self.n += 1


class _Dummy:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def step(self) -> None:
        pass


def get_profile_context(enable_flag: bool, warmup_steps: int, active_steps: int, save_dir: Optional[str] = None,
                        nsys: bool = False):
    if not enable_flag:
        return _Dummy()
    if nsys:
        return _NsysGate(warmup_steps, active_steps)
    from torch.profiler import ProfilerActivity, profile, schedule, tensorboard_trace_handler

    acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
    return profile(activities=acts, schedule=schedule(wait=0, warmup=warmup_steps, active=active_steps),
                   on_trace_ready=tensorboard_trace_handler(save_dir) if save_dir else None, record_shapes=True,
                   profile_memory=True, with_stack=True)


class PerformanceEvaluator:
    """tokens/s and model TFLOP/s per GPU: `on_step_start(step)` / `on_step_end(input_ids)` / `on_fit_end()`."""

    def __init__(self, model_numel: int, num_layers: int, hidden_size: int, vocab_size: int,
                 enable_grad_checkpoint: bool = False, ignore_steps: int = 0, dp_world_size: Optional[int] = None) -> None:
        self.model_numel, self.L, self.h, self.V = model_numel, num_layers, hidden_size, vocab_size
        self.ckpt, self.ignore_steps = enable_grad_checkpoint, ignore_steps
        self.dp_world_size = dp_world_size or (dist.get_world_size() if dist.is_initialized() else 1)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.mp_world_size = max(self.world // self.dp_world_size, 1)
        self.disable = False
        self._ev = None
        self._t0 = 0.0
        self.seconds = 0.0
        self.num_samples = 0
        self.flop_megatron = 0.0
        self.flop = 0.0

    def on_step_start(self, step: int) -> None:
        self.disable = self.ignore_steps > 0 and step < self.ignore_steps
        if self.disable:
            return
        if torch.cuda.is_available():
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._ev[0].record()
        else:
            self._t0 = time.perf_counter()

    def on_step_end(self, input_ids: torch.Tensor, **kwargs) -> None:
        if self.disable:
            return
        if torch.cuda.is_available():
            self._ev[1].record()
            torch.cuda.synchronize()
            self.seconds += self._ev[0].elapsed_time(self._ev[1]) / 1e3
        else:
            self.seconds += time.perf_counter() - self._t0
        B, S = input_ids.shape
        self.num_samples += B
        ck = 1 if self.ckpt else 0
        self.flop_megatron += 24 * (3 + ck) * B * S * self.L * self.h ** 2 * (
            1.0 + S / (6.0 * self.h) + self.V / (16.0 * self.L * self.h))
        self.flop += B * S * self.model_numel * 2 * (3 + ck)

    def on_fit_end(self) -> Dict[str, float]:
        t = torch.tensor([self.seconds], dtype=torch.float64)
        if dist.is_initialized():
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
            t = t.to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = max(float(t.item()), 1e-9)
        samples_s = self.num_samples * self.dp_world_size / sec
        out = {"seconds": sec, "samples_per_sec": samples_s,
               "tflops_per_gpu_megatron": self.flop_megatron / 1e12 / sec / self.mp_world_size,
               "tflops_per_gpu": self.flop / 1e12 / sec / self.mp_world_size}
        if not dist.is_initialized() or dist.get_rank() == 0:
            print(f"num_samples: {self.num_samples}, dp_world_size: {self.dp_world_size}, "
                  f"throughput: {samples_s:.2f} samples/s, TFLOPS/GPU (megatron): {out['tflops_per_gpu_megatron']:.2f}, "
                  f"TFLOPS/GPU: {out['tflops_per_gpu']:.2f}")
        return out


class CommProfiler:
    """Counts calls / bytes / time of `torch.distributed` collectives by monkey-patching them inside a context."""

    _OPS = ["all_reduce", "all_gather", "all_gather_into_tensor", "reduce_scatter", "reduce_scatter_tensor",
            "broadcast", "reduce", "all_to_all_single", "all_to_all", "send", "recv"]

    def __init__(self) -> None:
        self.stats: Dict[str, Dict[str, float]] = {}
        self._orig = {}

    @staticmethod
    def _nbytes(args, kwargs) -> int:
        n = 0
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a):
                n = max(n, a.numel() * a.element_size())
            elif isinstance(a, (list, tuple)) and a and torch.is_tensor(a[0]):
                n = max(n, sum(t.numel() * t.element_size() for t in a))
        return n

    def __enter__(self):
        for name in self._OPS:
            if not hasattr(dist, name):
                continue
            orig = getattr(dist, name)
            self._orig[name] = orig

            def wrapped(*args, __orig=orig, __name=name, **kwargs):
                t0 = time.perf_counter()
                out = __orig(*args, **kwargs)
                s = self.stats.setdefault(__name, {"count": 0, "bytes": 0, "host_seconds": 0.0})
                s["count"] += 1
                s["bytes"] += self._nbytes(args, kwargs)
                s["host_seconds"] += time.perf_counter() - t0
                return out

            setattr(dist, name, wrapped)
        return self

    def __exit__(self, *exc):
        for name, orig in self._orig.items():
            setattr(dist, name, orig)
        self._orig.clear()
        return False

    # legacy `ProfilerContext` protocol
    name, priority = "Comm", 5

    def enable(self) -> None:
        self.__enter__()

    def disable(self) -> None:
        self.__exit__(None, None, None)

    def show(self) -> None:
        print(self.result_str())

    def result_str(self, sep: str = "\n") -> str:
        lines = ["collective            calls        MiB   host s"]
        for k, s in sorted(self.stats.items(), key=lambda kv: -kv[1]["bytes"]):
            lines.append(f"{k:20s} {int(s['count']):6d} {s['bytes'] / 2**20:10.1f} {s['host_seconds']:8.3f}")
        return "\n".join(lines)
