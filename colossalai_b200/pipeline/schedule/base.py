"""Schedule base class.  Parity: reference `colossalai/pipeline/schedule/base.py`."""
from __future__ import annotations

from typing import Any, Callable, Iterable, Optional

import torch
from torch import Tensor
from torch.nn import Module

from ...accelerator import get_accelerator
from ...interface import OptimizerWrapper
from ..stage_manager import PipelineStageManager
from ._utils import get_batch_size, get_micro_batch, to_device
from torch.utils._pytree import tree_map

__all__ = ["PipelineSchedule"]


class PipelineSchedule:
    def __init__(self, stage_manager: PipelineStageManager) -> None:
        self.stage_manager = stage_manager

    def forward_backward_step(self, model: Module, data_iter: Iterable, criterion: Callable[[Any, Any], Tensor],
                              optimizer: Optional[OptimizerWrapper] = None, return_loss: bool = False,
                              return_outputs: bool = False) -> dict:
        raise NotImplementedError

    # ---- micro-batching shared by all schedules
    def load_batch(self, data_iter: Iterable, device: Optional[torch.device] = None) -> None:
        batch = next(data_iter) if not isinstance(data_iter, (dict, list, tuple)) else data_iter
        if device is None:
            device = get_accelerator().get_current_device()
        if device is not None:
            batch = tree_map(lambda x: to_device(x, device), batch)
        self.microbatch_offset = [0 for _ in range(getattr(self, "num_model_chunks", 1))]
        self.batch = batch
        self.batch_size = get_batch_size(batch)
        if self.microbatch_size is None:
            assert self.batch_size % self.num_microbatch == 0, "Batch size should divided by # microbatches"
            self.microbatch_size = self.batch_size // self.num_microbatch
        if self.num_microbatch is None:
            assert self.batch_size % self.microbatch_size == 0, "Batch size should divided by the microbatch size"
            self.num_microbatch = self.batch_size // self.microbatch_size
        if self.last_batch_size is None:
            self.last_batch_size = self.batch_size
        elif self.last_batch_size != self.batch_size:
            self.enable_metadata_cache = False    # shapes changed: resend metadata
            self.reset_metadata_cache()
            self.last_batch_size = self.batch_size

    def reset_metadata_cache(self) -> None:
        pass

    def load_micro_batch(self, model_chunk_id: int = 0) -> Any:
        off = self.microbatch_offset[model_chunk_id]
        mb = get_micro_batch(self.batch, off, self.microbatch_size)
        self.microbatch_offset[model_chunk_id] += self.microbatch_size
        return mb
