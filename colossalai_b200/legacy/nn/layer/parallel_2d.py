"""2D (SUMMA) tensor-parallel linear on a q x q mesh.
Parity: reference `colossalai/legacy/nn/layer/parallel_2d/{layers.py (Linear2D), _operation.py (Matmul_AB_2D :200,
Matmul_ABT_2D :352, Matmul_ATB_2D :510)}`.  The forward is SUMMA written with differentiable broadcasts, so autograd
produces the ABT / ATB passes of the reference's hand-written backward.

Layout: rank (i, j) holds X[i-th row block, j-th column block], W[i-th K block, j-th N block] and gets Y[i, j]."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...context import ParallelMode, global_context as gpc
from ._collectives import BroadcastFwdReduceBwd, IdentityFwdAllReduceBwd

__all__ = ["Linear2D", "split_2d", "matmul_ab_2d"]


def split_2d(x: torch.Tensor, row_dim: int = 0, col_dim: int = -1) -> torch.Tensor:
    """This rank's [i, j] block of a replicated tensor."""
    q = gpc.tensor_dims["q"]
    i = gpc.get_local_rank(ParallelMode.PARALLEL_2D_COL)       # row index = rank inside the column group
    j = gpc.get_local_rank(ParallelMode.PARALLEL_2D_ROW)       # column index = rank inside the row group
    return x.chunk(q, dim=row_dim)[i].chunk(q, dim=col_dim)[j].contiguous()


def matmul_ab_2d(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """C_ij = sum_k A_ik B_kj with A_ik broadcast along the row and B_kj along the column."""
    q = gpc.tensor_dims["q"]
    row_g, col_g = gpc.get_group(ParallelMode.PARALLEL_2D_ROW), gpc.get_group(ParallelMode.PARALLEL_2D_COL)
    out = None
    for k in range(q):
        a_k = BroadcastFwdReduceBwd.apply(a, k, row_g)        # owner: column k of my row
        b_k = BroadcastFwdReduceBwd.apply(b, k, col_g)        # owner: row k of my column
        part = a_k @ b_k
        out = part if out is None else out + part
    return out


class Linear2D(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None) -> None:
        super().__init__()
        q = gpc.tensor_dims["q"]
        assert in_features % q == 0 and out_features % q == 0
        self.in_features, self.out_features, self.q = in_features, out_features, q
        self.weight = nn.Parameter(torch.empty(in_features // q, out_features // q, dtype=dtype, device=device))
        # the bias block [N/q] is shared by the ranks of a column: keep it on every rank, sync its grad over the column
        self.bias = nn.Parameter(torch.zeros(out_features // q, dtype=dtype, device=device)) if bias else None
        nn.init.uniform_(self.weight, -1 / math.sqrt(in_features), 1 / math.sqrt(in_features))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: this rank's [..., M/q rows x K/q] block -> [..., N/q] block."""
        shape = x.shape
        y = matmul_ab_2d(x.reshape(-1, shape[-1]), self.weight)
        if self.bias is not None:
            y = y + IdentityFwdAllReduceBwd.apply(self.bias, gpc.get_group(ParallelMode.PARALLEL_2D_COL))
        return y.view(*shape[:-1], -1)
