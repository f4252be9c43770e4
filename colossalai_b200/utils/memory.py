"""Memory capacity helpers.  Parity: reference `colossalai/utils/memory.py:47`."""
from __future__ import annotations

import psutil
import torch


def colo_get_cpu_memory_capacity() -> int:
    return psutil.virtual_memory().total


def colo_device_memory_capacity(device: torch.device) -> int:
    device = torch.device(device)
    if device.type == "cpu":
        return colo_get_cpu_memory_capacity()
    if device.type == "cuda":
        return torch.cuda.get_device_properties(device).total_memory
    raise ValueError(device)


def colo_set_process_memory_fraction(ratio: float) -> None:
    if torch.cuda.is_available():
        torch.cuda.set_per_process_memory_fraction(ratio)


def bytes_to_GB(v: int, decimal: int = 2) -> float:
    return round(v / (1024**3), decimal)


def bytes_to_MB(v: int, decimal: int = 2) -> float:
    return round(v / (1024**2), decimal)
