"""3D tensor-parallel linear on a q x q x q cube.
Parity: reference `colossalai/legacy/nn/layer/parallel_3d/{layers.py (Linear3D), _operation.py (linear_3d)}`.

Ranks are (i, j, k).  X is stored as [M/(q*q) rows (split over i then k), K/q columns (split over j)], W as
[K/q rows (over j), N/(q*q) columns (split over k then i)], Y as [M/(q*q) rows (over i then j), N/q columns (over k)]:
    X_full_rows = all_gather(X, rows, over k)        (input group)
    W_full_cols = all_gather(W, cols, over i)        (weight group)
    Y = reduce_scatter(X_full_rows @ W_full_cols, rows, over j)   (output group)
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...context import ParallelMode, global_context as gpc
from ._collectives import AllGatherFwdReduceScatterBwd, ReduceScatterFwdAllGatherBwd

__all__ = ["Linear3D", "split_3d_input", "split_3d_weight", "gather_3d_output"]


def _coords():
    i = gpc.get_local_rank(ParallelMode.PARALLEL_3D_WEIGHT)
    j = gpc.get_local_rank(ParallelMode.PARALLEL_3D_INPUT)
    k = gpc.get_local_rank(ParallelMode.PARALLEL_3D_OUTPUT)
    return i, j, k


def split_3d_input(x: torch.Tensor) -> torch.Tensor:
    q = gpc.tensor_dims["q"]
    i, j, k = _coords()
    return x.chunk(q, dim=0)[i].chunk(q, dim=0)[k].chunk(q, dim=-1)[j].contiguous()


def split_3d_weight(w: torch.Tensor) -> torch.Tensor:
    """w: [K, N] replicated -> this rank's [K/q, N/q^2] block."""
    q = gpc.tensor_dims["q"]
    i, j, k = _coords()
    return w.chunk(q, dim=0)[j].chunk(q, dim=1)[k].chunk(q, dim=1)[i].contiguous()


def gather_3d_output(y: torch.Tensor) -> torch.Tensor:
    """Inverse of the output layout (tests / debugging): returns the full [M, N] on every rank."""
    from ....parallel import comm

    y = comm.all_gather(y.contiguous(), 0, gpc.get_group(ParallelMode.PARALLEL_3D_INPUT))    # rows over j
    y = comm.all_gather(y.contiguous(), 0, gpc.get_group(ParallelMode.PARALLEL_3D_WEIGHT))   # rows over i
    return comm.all_gather(y.contiguous(), 1, gpc.get_group(ParallelMode.PARALLEL_3D_OUTPUT))


class Linear3D(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = False, dtype=None, device=None) -> None:
        super().__init__()
        q = gpc.tensor_dims["q"]
        assert in_features % q == 0 and out_features % (q * q) == 0
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(in_features // q, out_features // (q * q), dtype=dtype, device=device))
        nn.init.uniform_(self.weight, -1 / math.sqrt(in_features), 1 / math.sqrt(in_features))
        assert not bias, "Linear3D: add the bias on the gathered output (kept out of the cube for clarity)"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xg = AllGatherFwdReduceScatterBwd.apply(x, 0, gpc.get_group(ParallelMode.PARALLEL_3D_OUTPUT))   # rows over k
        wg = AllGatherFwdReduceScatterBwd.apply(self.weight, 1, gpc.get_group(ParallelMode.PARALLEL_3D_WEIGHT))
        y = xg @ wg                                            # [M/q (rows of i), N/q (cols of k)] partial over j
        return ReduceScatterFwdAllGatherBwd.apply(y, 0, gpc.get_group(ParallelMode.PARALLEL_3D_INPUT))
