"""`MixedFusedLayerNorm` / `MixedFusedRMSNorm`: nn.Module faces of the native norm kernels (16-bit I/O, fp32 statistics).
Parity: reference `colossalai/nn/layer/layernorm.py:17-80` (`FusedLayerNormAffineFunction` over layernorm_cuda)."""
from __future__ import annotations

import numbers

import torch
import torch.nn as nn
from torch.nn import init

from ... import ops

__all__ = ["MixedFusedLayerNorm", "MixedFusedRMSNorm"]


class MixedFusedLayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps: float = 1e-5, device=None, dtype=None) -> None:
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = torch.Size(normalized_shape)
        self.eps = eps
        self.weight = nn.Parameter(torch.empty(*normalized_shape, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.empty(*normalized_shape, device=device, dtype=dtype))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        init.ones_(self.weight)
        init.zeros_(self.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = self.weight.numel()
        y = ops.layer_norm(x.reshape(-1, n), self.weight.reshape(-1), self.bias.reshape(-1), self.eps)
        return y.view(x.shape)

    def extra_repr(self) -> str:
        return f"{tuple(self.normalized_shape)}, eps={self.eps}"


class MixedFusedRMSNorm(nn.Module):
    def __init__(self, normalized_shape, eps: float = 1e-6, device=None, dtype=None) -> None:
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = torch.Size(normalized_shape)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(*normalized_shape, device=device, dtype=dtype))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = self.weight.numel()
        return ops.rms_norm(x.reshape(-1, n), self.weight.reshape(-1), self.eps).view(x.shape)
