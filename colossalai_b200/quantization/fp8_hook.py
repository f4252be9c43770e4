"""Rewrites `F.linear` on hooked parameters into the fp8 linear.
Parity: reference `colossalai/quantization/fp8_hook.py:7-23`."""
import torch.nn.functional as F

from ..tensor.param_op_hook import ColoParamOpHook
from .fp8 import linear_fp8

__all__ = ["FP8Hook"]


class FP8Hook(ColoParamOpHook):
    def pre_forward(self, params) -> None:
        pass

    def post_forward(self, params) -> None:
        pass

    def pre_backward(self, params) -> None:
        pass

    def post_backward(self, params) -> None:
        pass

    def rewrite_op(self, func):
        return linear_fp8 if func is F.linear else func


def convert_linear_to_fp8(module) -> None:
    """In place: every `nn.Linear` of `module` computes through `linear_fp8` (used by the plugins' `use_fp8` switch)."""
    import types

    import torch.nn as nn

    def fwd(self, x):
        return linear_fp8(x, self.weight, self.bias)

    for m in module.modules():
        if type(m) is nn.Linear:
            m.forward = types.MethodType(fwd, m)
