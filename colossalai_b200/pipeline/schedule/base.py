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


def _shape_signature(batch: Any, microbatch_size: int):
    """(shape beyond the batch dim, dtype) of every tensor of the batch + the micro-batch size."""
    from torch.utils._pytree import tree_flatten

    leaves, _ = tree_flatten(batch)
    return (microbatch_size,) + tuple((tuple(x.shape[1:]), str(x.dtype)) for x in leaves if torch.is_tensor(x))


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
        # the user fixed ONE of (number of micro-batches, micro-batch size); the other follows the batch - every batch,
        # not only the first one (a larger batch must not silently lose its tail, a smaller one must not over-read)
        if not hasattr(self, "_derive"):
            self._derive = "size" if self.microbatch_size is None else ("num" if self.num_microbatch is None else None)
        if self._derive == "size":
            assert self.batch_size % self.num_microbatch == 0, "Batch size should divided by # microbatches"
            self.microbatch_size = self.batch_size // self.num_microbatch
        elif self._derive == "num":
            assert self.batch_size % self.microbatch_size == 0, "Batch size should divided by the microbatch size"
            self.num_microbatch = self.batch_size // self.microbatch_size
        else:
            assert self.num_microbatch * self.microbatch_size == self.batch_size, (
                f"batch of {self.batch_size} != {self.num_microbatch} micro-batches x {self.microbatch_size}")
        # cached P2P metadata (shapes / dtypes of what a stage sends) is only valid while the micro-batches keep their
        # shape: every stage sees the same batch, so all of them drop the cache together and re-exchange it once
        sig = _shape_signature(batch, self.microbatch_size)
        if getattr(self, "_batch_signature", None) is not None and sig != self._batch_signature:
            self.reset_metadata_cache()
        self._batch_signature = sig
        self.last_batch_size = self.batch_size

    def reset_metadata_cache(self) -> None:
        pass

    def load_micro_batch(self, model_chunk_id: int = 0) -> Any:
        off = self.microbatch_offset[model_chunk_id]
        mb = get_micro_batch(self.batch, off, self.microbatch_size)
        self.microbatch_offset[model_chunk_id] += self.microbatch_size
        return mb
