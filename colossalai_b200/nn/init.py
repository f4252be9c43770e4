"""Initializer factories returning closures `(tensor, fan_in, fan_out)`.
Parity: reference `colossalai/nn/init.py:8-252`."""
from __future__ import annotations

import math
import warnings

import torch.nn as nn
from torch import Tensor

__all__ = ["zeros_", "ones_", "uniform_", "normal_", "trunc_normal_", "kaiming_uniform_", "kaiming_normal_",
           "xavier_uniform_", "xavier_normal_", "lecun_uniform_", "lecun_normal_"]


def zeros_():
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        return nn.init.zeros_(tensor)

    return initializer


def ones_():
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        return nn.init.ones_(tensor)

    return initializer


def uniform_(a: float = 0.0, b: float = 1.0):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        return nn.init.uniform_(tensor, a, b)

    return initializer


def normal_(mean: float = 0.0, std: float = 1.0):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        return nn.init.normal_(tensor, mean, std)

    return initializer


def trunc_normal_(mean: float = 0.0, std: float = 1.0, a: float = -2.0, b: float = 2.0):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        return nn.init.trunc_normal_(tensor, mean, std, a, b)

    return initializer


def _fan(mode: str, fan_in, fan_out) -> int:
    if mode == "fan_in":
        assert fan_in is not None, "Fan_in is not provided."
        return fan_in
    if mode == "fan_out":
        assert fan_out is not None, "Fan_out is not provided."
        return fan_out
    raise ValueError(f"Invalid initialization mode '{mode}'")


def kaiming_uniform_(a=0, mode="fan_in", nonlinearity="leaky_relu"):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        if 0 in tensor.shape:
            warnings.warn("Initializing zero-element tensors is a no-op")
            return tensor
        fan = _fan(mode, fan_in, fan_out)
        std = nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
        bound = math.sqrt(3.0) * std
        return nn.init.uniform_(tensor, -bound, bound)

    return initializer


def kaiming_normal_(a=0, mode="fan_in", nonlinearity="leaky_relu"):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        if 0 in tensor.shape:
            warnings.warn("Initializing zero-element tensors is a no-op")
            return tensor
        fan = _fan(mode, fan_in, fan_out)
        std = nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
        return nn.init.normal_(tensor, 0, std)

    return initializer


def xavier_uniform_(a: float = math.sqrt(3.0), scale: float = 2.0, gain: float = 1.0):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        assert fan_in is not None, "Fan_in is not provided."
        fan = fan_in + (fan_out if fan_out is not None else 0)
        std = gain * math.sqrt(scale / float(fan))
        bound = a * std
        return nn.init.uniform_(tensor, -bound, bound)

    return initializer


def xavier_normal_(scale: float = 2.0, gain: float = 1.0):
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        assert fan_in is not None, "Fan_in is not provided."
        fan = fan_in + (fan_out if fan_out is not None else 0)
        std = gain * math.sqrt(scale / float(fan))
        return nn.init.normal_(tensor, 0.0, std)

    return initializer


def lecun_uniform_():
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        assert fan_in is not None, "Fan_in is not provided."
        bound = math.sqrt(3.0 / fan_in)
        return nn.init.uniform_(tensor, -bound, bound)

    return initializer


def lecun_normal_():
    def initializer(tensor: Tensor, fan_in: int = None, fan_out: int = None):
        assert fan_in is not None, "Fan_in is not provided."
        std = math.sqrt(1.0 / fan_in)
        return nn.init.trunc_normal_(tensor, std=std / 0.87962566103423978)

    return initializer
