"""Chunk manager construction helper.  Parity: reference `colossalai/zero/gemini/chunk/utils.py`."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .manager import ChunkManager
from .search_utils import search_chunk_configuration

__all__ = ["init_chunk_manager", "safe_div"]


def safe_div(a, b):
    return 0 if a == 0 else a / b


def init_chunk_manager(model: nn.Module, init_device: Optional[torch.device] = None, hidden_dim: Optional[int] = None,
                       reuse_fp16_chunk: bool = True, verbose: bool = False, max_prefetch: int = 0,
                       **kwargs) -> ChunkManager:
    if hidden_dim:
        search_interval = hidden_dim
    else:
        search_interval = 1024
    kwargs["search_interval"] = search_interval
    dist.barrier() if dist.is_initialized() else None
    config, total, wasted = search_chunk_configuration(model, **kwargs)
    if verbose and (not dist.is_initialized() or dist.get_rank() == 0):
        mb = 1024**2
        print(f"searching chunk configuration: used {total / mb:.2f}M elements, wasted {wasted / mb:.2f}M "
              f"({100 * safe_div(wasted, total + wasted):.2f}%), chunk size {list(config.values())[0]['chunk_size']}",
              flush=True)
    return ChunkManager(config, init_device, reuse_fp16_chunk=reuse_fp16_chunk, max_prefetch=max_prefetch)
