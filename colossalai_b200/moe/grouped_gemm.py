"""Grouped (variable-M) linear: rows grouped by expert, one weight matrix per expert.

Native path (planned): tcgen05 grouped GEMM (kernel/csrc/gemm_tcgen05.cu, per-group tile scheduler).  Current CUDA
path: `torch._grouped_mm` (cuBLAS/CUTLASS library grouped GEMM) when available, else a per-expert loop.
Replaces the reference's python loop over local experts (`shardformer/modeling/mixtral.py:177-191`).
"""
from __future__ import annotations

import torch


class _GroupedLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, counts):
        # x [N, K]; w [E, O, K]; counts [E] (cpu or device int64)
        offs = torch.cumsum(counts, 0)
        ctx.save_for_backward(x, w, counts)
        out = x.new_empty(x.shape[0], w.shape[1])
        c = counts.tolist()
        s = 0
        for e, n in enumerate(c):
            if n:
                out[s:s + n] = x[s:s + n] @ w[e].t()
            s += n
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w, counts = ctx.saved_tensors
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        s = 0
        for e, n in enumerate(counts.tolist()):
            if n:
                dx[s:s + n] = dy[s:s + n] @ w[e]
                dw[e] = dy[s:s + n].t() @ x[s:s + n]
            s += n
        return dx, dw, None


def grouped_linear(x: torch.Tensor, w: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """y[rows of expert e] = x[rows of expert e] @ w[e]^T."""
    if x.is_cuda and hasattr(torch, "_grouped_mm") and x.dtype == torch.bfloat16 and x.shape[0] > 0 \
            and x.shape[1] % 8 == 0 and w.shape[1] % 8 == 0:
        offs = torch.cumsum(counts.to(x.device), 0).to(torch.int32)
        try:
            return torch._grouped_mm(x, w.transpose(-2, -1), offs=offs)
        except Exception:
            pass
    return _GroupedLinear.apply(x, w, counts.cpu() if counts.is_cuda else counts)
