"""Chunk-size search: pick the chunk size (in a range) that wastes the least memory when packing the model's params.
Parity: reference `colossalai/zero/gemini/chunk/search_utils.py:108-191`."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch.distributed as dist
import torch.nn as nn
from torch.distributed import ProcessGroup

from ....utils import is_ddp_ignored

__all__ = ["search_chunk_configuration", "classify_params_by_dp_degree", "in_ddp"]


def in_ddp(param: nn.Parameter) -> bool:
    return not is_ddp_ignored(param)


def _filter_exlarge_params(model: nn.Module, size_dict: Dict[int, List[int]]) -> None:
    """Ignore outliers (> mean + 3 std) when choosing the chunk size: they get a private chunk anyway."""
    agg = [s for sizes in size_dict.values() for s in sizes]
    if not agg:
        return
    arr = np.array(agg)
    upper = arr.mean() + 3 * arr.std()
    for k in size_dict:
        size_dict[k] = [s for s in size_dict[k] if s <= upper]


def _get_unused_byte(size_list: List[int], chunk_size: int) -> int:
    """Wasted elements if `size_list` is packed first-fit into chunks of `chunk_size`."""
    acc, left = 0, 0
    for s in size_list:
        if s > left:
            acc += left
            left = chunk_size
        left -= s
    return left + acc


def _tensor_numel(p: nn.Parameter) -> int:
    return p.numel()


def classify_params_by_dp_degree(param_order, process_group: Optional[ProcessGroup] = None) -> Dict[int, List[nn.Parameter]]:
    """All params of a ZeRO group share one data-parallel degree here (TP-sharded params already hold local shards)."""
    ws = dist.get_world_size(process_group) if dist.is_initialized() else 1
    out: Dict[int, List[nn.Parameter]] = {ws: []}
    params = param_order.generate() if hasattr(param_order, "generate") else param_order
    for p in params:
        if in_ddp(p):
            out[ws].append(p)
    return out


def search_chunk_configuration(model: nn.Module, search_range_m: float, search_interval: int = 1024,
                               min_chunk_size_m: float = 32, filter_exlarge_params: bool = True,
                               strict_ddp_flag: bool = False, process_group: Optional[ProcessGroup] = None,
                               memstas=None) -> Tuple[Dict, int, int]:
    """Returns (config {dp_degree: {chunk_size, keep_gathered}}, total elements, wasted elements)."""
    search_range = round(search_range_m * 1024**2)
    min_chunk_size = round(min_chunk_size_m * 1024**2)
    assert search_range >= 0
    params_dict = classify_params_by_dp_degree(list(model.parameters()), process_group)
    size_lcm = np.lcm.reduce(list(params_dict.keys()))
    config_dict: Dict[int, Dict] = {}
    total_param_size = 0
    size_dict: Dict[int, List[int]] = {}
    for dp_degree, plist in params_dict.items():
        sizes = [_tensor_numel(p) for p in plist]
        total_param_size += sum(sizes)
        size_dict[dp_degree] = sizes
    if filter_exlarge_params:
        _filter_exlarge_params(model, size_dict)
    max_size = max([max(v) for v in size_dict.values() if v] or [min_chunk_size])
    start_size = max(min_chunk_size, max_size)
    start_size = int(math.ceil(start_size / search_interval) * search_interval)
    min_waste, best = float("+inf"), start_size
    for chunk_size in range(start_size, start_size + search_range + 1, search_interval):
        waste = sum(_get_unused_byte(sizes, chunk_size) for sizes in size_dict.values())
        if waste < min_waste:
            min_waste, best = waste, chunk_size
    best = best + (-best % size_lcm)
    for dp_degree in params_dict:
        config_dict[dp_degree] = dict(chunk_size=int(best), keep_gathered=False)
    return config_dict, total_param_size, int(min_waste)
