"""Moves a model to a device when the plugin does not control placement.
Parity: reference `colossalai/booster/accelerator.py:16-54`."""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ["Accelerator"]

_supported_devices = ["cpu", "cuda"]


class Accelerator:
    def __init__(self, device: str) -> None:
        self.device = device
        assert device in _supported_devices, f"device must be one of {_supported_devices}, got {device}"

    def bind(self) -> None:
        if self.device == "cuda" and torch.cuda.is_available():
            import os

            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)) % torch.cuda.device_count())

    def configure_model(self, model: nn.Module) -> nn.Module:
        return model.to(torch.device(self.device))
