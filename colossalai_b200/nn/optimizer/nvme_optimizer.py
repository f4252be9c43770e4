"""Base class that can park optimizer states on NVMe between steps.

Parity: reference `colossalai/nn/optimizer/nvme_optimizer.py:10-167` (tensornvme DiskOffloader).  Our offloader is
the native async file IO library (`kernel/csrc/async_file_io.cpp`, pthread workers + pread/pwrite) — no tensornvme.
"""
from __future__ import annotations

import math
import os
import tempfile
from typing import Callable, Dict, List, Optional

import torch
from torch.nn.parameter import Parameter

__all__ = ["NVMeOptimizer"]


class NVMeOptimizer(torch.optim.Optimizer):
    """`nvme_offload_fraction` of the parameters (by element count) keep their states on disk; states are prefetched
    one parameter ahead while the current parameter is being updated."""

    num_fp32_shards_per_param = 0

    def __init__(self, params, defaults: dict, nvme_offload_fraction: float = 0.0,
                 offload_dir: Optional[str] = None) -> None:
        assert 0.0 <= nvme_offload_fraction <= 1.0
        super().__init__(params, defaults)
        self.nvme_offload_fraction = float(nvme_offload_fraction)
        self.offloader = None
        self.is_on_nvme: Dict[Parameter, bool] = {}
        self.offloaded_numel = 0
        self.total_numel: Optional[int] = None
        self.can_offload_numel: Optional[int] = None
        self.prefetch_params: List[Parameter] = []
        self.param_to_prefetch_idx: Dict[Parameter, int] = {}
        if self.nvme_offload_fraction > 0.0:
            from ...utils.aio import DiskOffloader

            self.offload_dir = offload_dir or tempfile.mkdtemp(prefix="cb200_nvme_")
            self.offloader = DiskOffloader(self.offload_dir, n_entries=8)

    def _get_numel(self) -> int:
        return sum(p.numel() for g in self.param_groups for p in g["params"])

    def _post_state_init(self, param: Parameter) -> None:
        if self.offloader is None:
            return
        if self.total_numel is None:
            self.total_numel = self._get_numel()
            self.can_offload_numel = math.floor(self.total_numel * self.nvme_offload_fraction)
        numel = param.numel()
        if param.device.type == "cpu" and numel + self.offloaded_numel <= self.can_offload_numel:
            self.is_on_nvme[param] = True
            self.offloaded_numel += numel
        else:
            self.is_on_nvme[param] = False

    def _setup_prefetch_params(self) -> None:
        if self.offloader is None:
            return
        self.prefetch_params, self.param_to_prefetch_idx = [], {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if len(self.state[p]) > 0 and self.is_on_nvme.get(p, False):
                    self.param_to_prefetch_idx[p] = len(self.prefetch_params)
                    self.prefetch_params.append(p)

    def _pre_step(self, *state_keys: str) -> None:
        self._setup_prefetch_params()
        if self.offloader is None or not self.prefetch_params:
            return
        st = self.state[self.prefetch_params[0]]
        for k in state_keys:
            self.offloader.async_read(st[k])

    def _pre_update(self, param: Parameter, *state_keys: str) -> None:
        if self.offloader is None or param not in self.param_to_prefetch_idx:
            return
        self.offloader.sync_read_events()
        idx = self.param_to_prefetch_idx[param]
        if idx + 1 < len(self.prefetch_params):
            st = self.state[self.prefetch_params[idx + 1]]
            for k in state_keys:
                self.offloader.async_read(st[k])

    def _post_update(self, param: Parameter, *state_keys: str) -> None:
        if self.offloader is None:
            return
        self.offloader.sync_write_events()
        if self.is_on_nvme.get(param, False):
            st = self.state[param]
            for k in state_keys:
                self.offloader.async_write(st[k])

    def _post_step(self) -> None:
        if self.offloader is not None:
            self.offloader.synchronize()

    def step(self, closure: Optional[Callable[[], float]] = None):
        raise NotImplementedError

    def state_dict(self) -> dict:
        if self.offloader is not None:
            raise NotImplementedError("call load/flush before checkpointing an NVMe-offloaded optimizer")
        return super().state_dict()
