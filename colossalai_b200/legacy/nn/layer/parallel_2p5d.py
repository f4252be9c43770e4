"""2.5D tensor-parallel linear on a d x q x q mesh.
Parity: reference `colossalai/legacy/nn/layer/parallel_2p5d/{layers.py (Linear2p5D), _operation.py (Matmul_AB_2p5D)}`:
the batch is additionally split over the depth axis, every depth layer runs 2D SUMMA with the SAME weight blocks, and
weight gradients are summed over depth."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...context import ParallelMode, global_context as gpc
from ._collectives import BroadcastFwdReduceBwd, IdentityFwdAllReduceBwd

__all__ = ["Linear2p5D", "split_2p5d"]


def split_2p5d(x: torch.Tensor, row_dim: int = 0, col_dim: int = -1) -> torch.Tensor:
    """Rows are split over depth first, then over the mesh rows; columns over the mesh columns."""
    q, d = gpc.tensor_dims["q"], gpc.tensor_dims["d"]
    dep = gpc.get_local_rank(ParallelMode.PARALLEL_2P5D_DEP)
    i = gpc.get_local_rank(ParallelMode.PARALLEL_2P5D_COL)
    j = gpc.get_local_rank(ParallelMode.PARALLEL_2P5D_ROW)
    return x.chunk(d, dim=row_dim)[dep].chunk(q, dim=row_dim)[i].chunk(q, dim=col_dim)[j].contiguous()


class Linear2p5D(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None) -> None:
        super().__init__()
        q = gpc.tensor_dims["q"]
        assert in_features % q == 0 and out_features % q == 0
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(in_features // q, out_features // q, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(out_features // q, dtype=dtype, device=device)) if bias else None
        nn.init.uniform_(self.weight, -1 / math.sqrt(in_features), 1 / math.sqrt(in_features))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        q = gpc.tensor_dims["q"]
        row_g = gpc.get_group(ParallelMode.PARALLEL_2P5D_ROW)
        col_g = gpc.get_group(ParallelMode.PARALLEL_2P5D_COL)
        dep_g = gpc.get_group(ParallelMode.PARALLEL_2P5D_DEP)
        shape = x.shape
        a = x.reshape(-1, shape[-1])
        w = IdentityFwdAllReduceBwd.apply(self.weight, dep_g)     # same block on every depth layer: sum its grads
        out = None
        for k in range(q):
            part = BroadcastFwdReduceBwd.apply(a, k, row_g) @ BroadcastFwdReduceBwd.apply(w, k, col_g)
            out = part if out is None else out + part
        if self.bias is not None:
            b = IdentityFwdAllReduceBwd.apply(IdentityFwdAllReduceBwd.apply(self.bias, col_g), dep_g)
            out = out + b
        return out.view(*shape[:-1], -1)
