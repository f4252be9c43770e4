"""Inference helpers.  Parity: reference `colossalai/inference/utils.py` (ports, alibi slopes, model size,
checkpoint index detection)."""
from __future__ import annotations

import math
import socket
from contextlib import closing
from pathlib import Path
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

__all__ = ["find_available_ports", "get_alibi_slopes", "get_model_size", "has_index_file", "can_use_flash_attn2"]


def find_available_ports(num: int) -> List[int]:
    socks, ports = [], []
    for _ in range(num):
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


def get_alibi_slopes(num_heads: int, device: torch.device = None) -> torch.Tensor:
    """Geometric ALiBi slopes, with the standard interleave extension for non-power-of-two head counts."""
    p2 = 2 ** math.floor(math.log2(num_heads))
    base = 2.0 ** (-(2.0 ** -(math.log2(p2) - 3)))
    slopes = [base ** (i + 1) for i in range(p2)]
    if p2 != num_heads:
        extra_base = 2.0 ** (-(2.0 ** -(math.log2(2 * p2) - 3)))
        slopes += [extra_base ** (2 * i + 1) for i in range(num_heads - p2)]
    return torch.tensor(slopes, dtype=torch.float32, device=device)


def get_model_size(model: nn.Module) -> float:
    """Model size in GB (parameters + buffers)."""
    total = sum(p.numel() * p.element_size() for p in model.parameters())
    total += sum(b.numel() * b.element_size() for b in model.buffers())
    return total / 1024 ** 3


def has_index_file(checkpoint_path: str) -> Tuple[bool, Optional[Path]]:
    p = Path(checkpoint_path)
    if p.is_file():
        return (p.name.endswith(".index.json"), p if p.name.endswith(".index.json") else None)
    if p.is_dir():
        idx = sorted(p.glob("*.index.*json"))
        if len(idx) == 1:
            return True, idx[0]
        if len(idx) > 1:
            raise ValueError(f"Found multiple index files in {checkpoint_path}")
    return False, None


def can_use_flash_attn2(dtype: torch.dtype) -> bool:
    return dtype in (torch.float16, torch.bfloat16) and torch.cuda.is_available()
