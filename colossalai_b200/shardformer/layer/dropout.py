"""Dropout with TP-aware RNG.  Parity: reference `colossalai/shardformer/layer/dropout.py:13,50`."""
from __future__ import annotations

from typing import List, Union

import torch
import torch.nn as nn
from torch.distributed import ProcessGroup

from .parallel_module import ParallelModule
from .utils import create_randomizer_with_offset

__all__ = ["DropoutForParallelInput", "DropoutForReplicatedInput"]


class DropoutForParallelInput(ParallelModule, nn.Dropout):
    """Input is sharded over TP: every rank must draw a DIFFERENT mask (per-rank RNG stream)."""

    def __init__(self, p: float = 0.5, inplace: bool = False, process_group: ProcessGroup = None) -> None:
        nn.Module.__init__(self)
        self.p, self.inplace = p, inplace
        self.randomizer = create_randomizer_with_offset(torch.initial_seed() % (2**31), process_group)

    @staticmethod
    def from_native_module(module: nn.Dropout, process_group: Union[ProcessGroup, List[ProcessGroup]] = None, **kw):
        return DropoutForParallelInput(module.p, module.inplace, process_group)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.training or self.p == 0.0:
            return x
        with self.randomizer.fork_rng(enable_cpu=x.device.type == "cpu"):
            return nn.functional.dropout(x, self.p, True, self.inplace)


class DropoutForReplicatedInput(ParallelModule, nn.Dropout):
    """Input is replicated over TP: every rank must draw the SAME mask (shared RNG stream)."""

    def __init__(self, p: float = 0.5, inplace: bool = False, process_group: ProcessGroup = None) -> None:
        nn.Module.__init__(self)
        self.p, self.inplace = p, inplace
        self.randomizer = create_randomizer_with_offset(torch.initial_seed() % (2**31), process_group,
                                                        offset_by_rank=False)

    @staticmethod
    def from_native_module(module: nn.Dropout, process_group: Union[ProcessGroup, List[ProcessGroup]] = None, **kw):
        return DropoutForReplicatedInput(module.p, module.inplace, process_group)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.training or self.p == 0.0:
            return x
        with self.randomizer.fork_rng(enable_cpu=x.device.type == "cpu"):
            return nn.functional.dropout(x, self.p, True, self.inplace)
