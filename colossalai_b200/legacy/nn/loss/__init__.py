"""Tensor-parallel aware losses.  Parity: reference `colossalai/legacy/nn/loss/{__init__.py, loss_1d.py:1-110}`
(`CrossEntropyLoss` dispatching on the tensor-parallel mode, `VocabParallelCrossEntropyLoss1D`)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ....shardformer.layer.loss import cross_entropy_1d
from ...context import ParallelMode, global_context as gpc

__all__ = ["CrossEntropyLoss", "VocabParallelCrossEntropyLoss1D"]


class VocabParallelCrossEntropyLoss1D(nn.Module):
    """Logits sharded on the class dimension over `ParallelMode.PARALLEL_1D`."""

    def __init__(self, reduction: bool = True) -> None:
        super().__init__()
        self.reduction_mean = reduction

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        g = gpc.get_group(ParallelMode.PARALLEL_1D) if gpc.is_initialized(ParallelMode.PARALLEL_1D) else None
        return cross_entropy_1d(logits.reshape(-1, logits.shape[-1]), targets.reshape(-1), process_group=g,
                                mode="mean" if self.reduction_mean else "sum")


class CrossEntropyLoss(nn.Module):
    """Plain CE when logits are replicated, vocab-parallel CE when the classifier output stays sharded."""

    def __init__(self, reduction: bool = True, vocab_parallel: bool = False, **kw) -> None:
        super().__init__()
        self.vocab_parallel = vocab_parallel
        self.inner = VocabParallelCrossEntropyLoss1D(reduction) if vocab_parallel else None
        self.reduction = "mean" if reduction else "sum"
        self.kw = kw

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if self.inner is not None:
            return self.inner(logits, targets)
        return F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), targets.reshape(-1),
                               reduction=self.reduction, **self.kw)
