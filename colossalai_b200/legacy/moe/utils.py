"""Noisy-gating helpers.  Parity: reference `colossalai/legacy/moe/utils.py:1-120`."""
from __future__ import annotations

import torch


class NormalNoiseGenerator:
    """logits + N(0, 1/E^2) (Switch / GShard style jitter)."""

    def __init__(self, num_experts: int) -> None:
        self.std = 1.0 / num_experts ** 2

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return inputs + torch.randn_like(inputs) * self.std


class UniformNoiseGenerator:
    """logits * U(1 - eps, 1 + eps)."""

    def __init__(self, eps: float = 1e-2) -> None:
        self.eps = eps

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return inputs * torch.empty_like(inputs).uniform_(1.0 - self.eps, 1.0 + self.eps)


def get_noise_generator(noise_type, num_experts: int):
    if noise_type is None:
        return None
    if noise_type == "Jitter":
        return UniformNoiseGenerator()
    if noise_type == "Gaussian":
        return NormalNoiseGenerator(num_experts)
    raise NotImplementedError(f"unsupported input noise {noise_type}")
