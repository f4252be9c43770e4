"""Memory statistics used by the auto placement policy.

Parity: reference `colossalai/zero/gemini/memory_tracer/{memory_stats,memstats_collector,chunk_memstats_collector,
memory_monitor,runtime_mem_tracer,utils}.py`: sample the allocator's peak memory at every parameter-use boundary of
the warm-up iteration; non-model data = overall peak - chunk (model data) memory per period.
"""
from __future__ import annotations

import time
from typing import Any, List, Optional

import torch

from ....accelerator import get_accelerator

__all__ = ["MemStats", "MemStatsCollector", "ChunkMemStatsCollector", "SyncCudaMemoryMonitor", "AsyncMemoryMonitor",
           "RuntimeMemTracer", "colo_model_data_tensor_move", "get_cuda_memory_used"]


def get_cuda_memory_used() -> int:
    return torch.cuda.memory_allocated() if torch.cuda.is_available() else 0


class SyncCudaMemoryMonitor:
    """Peak memory between `start()` and `finish()` (synchronises the device at both ends)."""

    def __init__(self, power: int = 10) -> None:
        self.time_stamps: List[float] = []
        self.mem_stats: List[int] = []

    def __len__(self) -> int:
        return len(self.mem_stats)

    def start(self) -> None:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()

    def finish(self) -> int:
        if not torch.cuda.is_available():
            self.mem_stats.append(0)
            return 0
        torch.cuda.synchronize()
        self.time_stamps.append(time.time())
        peak = torch.cuda.max_memory_allocated()
        self.mem_stats.append(peak)
        return peak

    def clear(self) -> None:
        self.mem_stats.clear()
        self.time_stamps.clear()


class AsyncMemoryMonitor(SyncCudaMemoryMonitor):
    """API-compatible alias: sampling happens at op boundaries, no background thread is needed with CUDA's own
    peak counters."""


class MemStats:
    def __init__(self) -> None:
        self._step_param_dict = {}
        self._param_step_dict = {}
        self._step_nmd_dict = {}
        self._param_runtime_order = []
        self._preop_step = 0
        self._prev_overall_cuda = -1
        self._max_overall_cuda = 0
        self._prev_md_cuda = -1
        self._non_model_data_cuda_list: List[int] = []
        self._non_model_data_cpu_list: List[int] = []

    def calc_max_cuda_non_model_data(self) -> None:
        if self._prev_overall_cuda != -1 and self._prev_md_cuda != -1:
            nmd = max(self._prev_overall_cuda - self._prev_md_cuda, 0)
            self._step_nmd_dict[self._preop_step] = nmd
            self._non_model_data_cuda_list.append(nmd)

    def record_max_cuda_model_data(self, val: int) -> None:
        self._prev_md_cuda = val

    def record_max_cuda_overall_data(self, val: int) -> None:
        self._prev_overall_cuda = val
        self._max_overall_cuda = max(self._max_overall_cuda, val)

    @property
    def max_overall_cuda(self) -> int:
        return self._max_overall_cuda

    def increase_preop_step(self, param_list: List[torch.nn.Parameter]) -> None:
        for p in param_list:
            if p not in self._param_step_dict:
                self._param_step_dict[p] = [self._preop_step]
            else:
                self._param_step_dict[p].append(self._preop_step)
            self._param_runtime_order.append(p)
        self._step_param_dict[self._preop_step] = param_list
        self._preop_step += 1

    def param_used_step(self, param) -> Optional[List[int]]:
        return self._param_step_dict.get(param)

    def param_order(self):
        return self._param_runtime_order

    def non_model_data_list(self, device_type: str) -> List[int]:
        return self._non_model_data_cuda_list if device_type != "cpu" else self._non_model_data_cpu_list

    def max_non_model_data(self, device_type: str) -> float:
        lst = self.non_model_data_list(device_type)
        return max(lst) if lst else 0

    def clear(self) -> None:
        self.__init__()


class MemStatsCollector:
    def __init__(self) -> None:
        self._mem_monitor = SyncCudaMemoryMonitor()
        self._sampling_time: List[float] = []
        self._start_flag = False
        self._step_idx = 0
        self._step_total = 0
        self._memstats = MemStats()

    def next_period_non_model_data_usage(self, device_type: str) -> int:
        assert not self._start_flag, "Cannot get mem stats info during collection phase."
        assert self._step_total > 0, "Cannot get mem stats info before collection phase."
        lst = self._memstats.non_model_data_list(device_type)
        nxt = lst[self._step_idx] if self._step_idx < len(lst) else (lst[-1] if lst else 0)
        self._step_idx = (self._step_idx + 1) % max(self._step_total, 1)
        return nxt

    @property
    def sampling_time(self):
        return [t - self._sampling_time[0] for t in self._sampling_time]

    def start_collection(self) -> None:
        self._start_flag = True
        self._mem_monitor.start()

    def finish_collection(self) -> None:
        self.sample_overall_data()
        self._step_total = len(self._memstats.non_model_data_list("cuda"))
        self._start_flag = False

    def record_model_data_volume(self) -> None:
        raise NotImplementedError("use ChunkMemStatsCollector")

    def sample_overall_data(self) -> None:
        if self._start_flag:
            cuda_overall = self._mem_monitor.finish()
            self._memstats.record_max_cuda_overall_data(cuda_overall)
            self._memstats.calc_max_cuda_non_model_data()
            self._mem_monitor.start()
            self._sampling_time.append(time.time())

    def clear(self) -> None:
        self._memstats.clear()
        self._start_flag = False
        self._step_idx = 0
        self._step_total = 0


class ChunkMemStatsCollector(MemStatsCollector):
    def __init__(self, chunk_manager, memstats: Optional[MemStats] = None) -> None:
        super().__init__()
        self._chunk_manager = chunk_manager
        if memstats is not None:
            self.use_outside_memstats = True
            self._memstats = memstats
        else:
            self.use_outside_memstats = False

    def record_model_data_volume(self) -> None:
        if self._start_flag and not self.use_outside_memstats:
            self._memstats.record_max_cuda_model_data(self._chunk_manager.total_mem["cuda"])

    @property
    def cuda_margin_mem(self) -> float:
        total = get_accelerator().mem_get_info()[1] if torch.cuda.is_available() else 0
        return total - self._memstats.max_overall_cuda


def colo_model_data_tensor_move(src_t: torch.Tensor, tgt_t: torch.Tensor) -> None:
    tgt_t.data.copy_(src_t.data)
    src_t.data = torch.empty(0, device=src_t.device, dtype=src_t.dtype)


class RuntimeMemTracer:
    """Run one fwd+bwd of a module and record peak non-model memory per parameter use (static placement helper)."""

    def __init__(self, module: torch.nn.Module, dtype: torch.dtype = torch.half) -> None:
        self.module = module
        self.dtype = dtype
        self._memstats = MemStats()
        self._monitor = SyncCudaMemoryMonitor()
        self._hooks = []

    def parameters_in_runtime_order(self):
        return self._memstats._param_runtime_order

    def memstats(self) -> MemStats:
        return self._memstats

    def __call__(self, *args, **kwargs):
        def pre(mod, inp):
            ps = list(mod.parameters(recurse=False))
            if ps:
                self._memstats.record_max_cuda_overall_data(self._monitor.finish())
                self._memstats.record_max_cuda_model_data(sum(p.numel() * p.element_size() for p in self.module.parameters()))
                self._memstats.calc_max_cuda_non_model_data()
                self._memstats.increase_preop_step(ps)
                self._monitor.start()

        for m in self.module.modules():
            self._hooks.append(m.register_forward_pre_hook(pre))
        self._monitor.start()
        try:
            out = self.module(*args, **kwargs)
        finally:
            for h in self._hooks:
                h.remove()
            self._hooks.clear()
        return out
