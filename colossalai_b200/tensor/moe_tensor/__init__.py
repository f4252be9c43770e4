"""Tag expert parameters with their expert-parallel group.
Parity: reference `colossalai/tensor/moe_tensor/api.py:10-153`."""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup

__all__ = ["is_moe_tensor", "set_moe_tensor_ep_group", "get_ep_group", "get_ep_size", "get_ep_rank", "get_dp_group",
           "get_dp_size", "get_dp_rank", "set_moe_tensor_info"]


def is_moe_tensor(t: torch.Tensor) -> bool:
    return hasattr(t, "ep_group")


def set_moe_tensor_ep_group(t: torch.Tensor, ep_group: ProcessGroup, moe_dp_group: ProcessGroup = None) -> None:
    t.__setattr__("ep_group", ep_group)
    if moe_dp_group is not None:
        t.__setattr__("moe_dp_group", moe_dp_group)


set_moe_tensor_info = set_moe_tensor_ep_group


def get_ep_group(t: torch.Tensor) -> ProcessGroup:
    return t.ep_group


def get_ep_size(t: torch.Tensor) -> int:
    return dist.get_world_size(t.ep_group) if dist.is_initialized() else 1


def get_ep_rank(t: torch.Tensor) -> int:
    return dist.get_rank(t.ep_group) if dist.is_initialized() else 0


def get_dp_group(t: torch.Tensor) -> ProcessGroup:
    return getattr(t, "moe_dp_group", None)


def get_dp_size(t: torch.Tensor) -> int:
    g = get_dp_group(t)
    return dist.get_world_size(g) if (dist.is_initialized() and g is not None) else 1


def get_dp_rank(t: torch.Tensor) -> int:
    g = get_dp_group(t)
    return dist.get_rank(g) if (dist.is_initialized() and g is not None) else 0
