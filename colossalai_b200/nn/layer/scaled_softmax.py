"""`FusedScaleMaskSoftmax`: scale + mask + softmax in one kernel (padding or causal masks).
Parity: reference `colossalai/nn/layer/scaled_softmax.py:17-190` (`AttnMaskType`, `ScaledUpperTriangMaskedSoftmax`,
`ScaledMaskedSoftmax`, `FusedScaleMaskSoftmax` with its `is_kernel_available` / torch fallback split).  The native
kernels here have no 2048-key limit, so the fallback only triggers for fp32 inputs with `softmax_in_fp32` off."""
from __future__ import annotations

import enum
from typing import Callable, Optional

import torch
import torch.nn as nn

from ...ops.softmax import scaled_causal_softmax, scaled_masked_softmax

__all__ = ["AttnMaskType", "FusedScaleMaskSoftmax"]


class AttnMaskType(enum.Enum):
    padding = 1
    causal = 2
    paddedcausal = 3


class FusedScaleMaskSoftmax(nn.Module):
    def __init__(self, input_in_fp16: bool = False, input_in_bf16: bool = False,
                 attn_mask_type: AttnMaskType = AttnMaskType.padding, scaled_masked_softmax_fusion: bool = True,
                 mask_func: Optional[Callable] = None, softmax_in_fp32: bool = True, scale: Optional[float] = None):
        super().__init__()
        assert not (input_in_fp16 and input_in_bf16), "both fp16 and bf16 flags cannot be active at the same time."
        self.input_in_float16 = input_in_fp16 or input_in_bf16
        self.attn_mask_type = attn_mask_type
        self.scaled_masked_softmax_fusion = scaled_masked_softmax_fusion
        self.mask_func = mask_func
        self.softmax_in_fp32 = softmax_in_fp32
        self.scale = scale
        assert self.scale is None or softmax_in_fp32, "softmax should be in fp32 when scaled"

    def is_kernel_available(self, mask, b, np, sq, sk) -> bool:
        return bool(self.scaled_masked_softmax_fusion) and sk > 0

    def forward(self, input: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        assert input.dim() == 4, "expected [b, np, sq, sk]"
        b, np_, sq, sk = input.shape
        scale = self.scale if self.scale is not None else 1.0
        if self.is_kernel_available(mask, b, np_, sq, sk):
            if self.attn_mask_type == AttnMaskType.causal:
                return scaled_causal_softmax(input.view(-1, sq, sk), scale).view(b, np_, sq, sk)
            if self.attn_mask_type == AttnMaskType.paddedcausal and mask is not None:
                causal = torch.ones(sq, sk, dtype=torch.bool, device=input.device).triu(diagonal=sk - sq + 1)
                mask = mask.bool() | causal
            return scaled_masked_softmax(input, mask, scale)
        return self.forward_torch_softmax(input, mask)

    def forward_torch_softmax(self, input: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        x = input.float() if (self.input_in_float16 and self.softmax_in_fp32) else input
        if self.scale is not None:
            x = x * self.scale
        if mask is not None:
            x = self.mask_func(x, mask) if self.mask_func is not None else x.masked_fill(mask.bool(), -10000.0)
        p = torch.softmax(x, dim=-1)
        return p.to(input.dtype)
