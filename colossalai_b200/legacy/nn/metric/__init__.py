"""Accuracy metric aware of class-sharded logits.  Parity: reference `colossalai/legacy/nn/metric/*.py`."""
from __future__ import annotations

import torch
import torch.nn as nn

from ....parallel import comm
from ...context import ParallelMode, global_context as gpc

__all__ = ["Accuracy", "calc_acc"]


def calc_acc(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    return (logits.argmax(dim=-1) == targets).sum()


class Accuracy(nn.Module):
    def __init__(self, vocab_parallel: bool = False) -> None:
        super().__init__()
        self.vocab_parallel = vocab_parallel

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if self.vocab_parallel and gpc.is_initialized(ParallelMode.PARALLEL_1D):
            logits = comm.all_gather(logits.contiguous(), -1, gpc.get_group(ParallelMode.PARALLEL_1D))
        return calc_acc(logits, targets)
