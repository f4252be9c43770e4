"""Fused bias / activation / dropout / residual ops.

Parity: reference `colossalai/kernel/jit/{bias_dropout_add,bias_gelu,option}.py` (TorchScript fusions used by the
OPT / T5 / ChatGLM / BLIP2 / BERT / ViT policies with `enable_jit_fused`).  Here bias+activation is one native kernel
(`ops.bias_act`, kernel/csrc/elementwise.cu); dropout keeps torch's RNG, so bias+dropout+residual is expressed with
torch ops (one extra pass only in training with p > 0).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ...ops import bias_act

__all__ = ["bias_gelu", "bias_gelu_impl", "bias_dropout_add", "bias_dropout_add_fused_train",
           "bias_dropout_add_fused_inference", "set_jit_fusion_options"]


def bias_gelu(bias: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return bias_act(y, bias, "gelu_tanh")


bias_gelu_impl = bias_gelu


def bias_dropout_add(x: torch.Tensor, bias: torch.Tensor, residual: torch.Tensor, prob: float, training: bool
                     ) -> torch.Tensor:
    out = x if bias is None else x + bias
    if training and prob > 0:
        out = F.dropout(out, p=prob, training=True)
    return residual + out


def bias_dropout_add_fused_train(x, bias, residual, prob: float) -> torch.Tensor:
    return bias_dropout_add(x, bias, residual, prob, True)


def bias_dropout_add_fused_inference(x, bias, residual, prob: float) -> torch.Tensor:
    return bias_dropout_add(x, bias, residual, prob, False)


def set_jit_fusion_options() -> None:
    """No-op: there is no TorchScript fuser to configure (kept for API parity with `kernel/jit/option.py`)."""
