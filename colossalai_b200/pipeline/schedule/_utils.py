"""Pytree helpers for pipeline schedules.  Parity: reference `colossalai/pipeline/schedule/_utils.py`."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import torch
from torch.utils._pytree import tree_flatten, tree_map, tree_unflatten

__all__ = ["to_device", "get_batch_size", "get_micro_batch", "model_forward", "retain_grad", "require_grad", "detach",
           "clone", "release_tensor_data", "merge_batch", "tree_map_hf", "deallocate"]


def tree_map_hf(fn, pytree):
    return tree_map(fn, dict(pytree) if hasattr(pytree, "keys") and not isinstance(pytree, dict) else pytree)


def to_device(x: Any, device: Optional[torch.device] = None) -> Any:
    return x.to(device) if isinstance(x, torch.Tensor) else x


def get_batch_size(batch: Any) -> int:
    for leaf in tree_flatten(batch)[0]:
        if isinstance(leaf, torch.Tensor):
            return leaf.size(0)
    raise RuntimeError("No tensor found in the batch")


def get_micro_batch(batch: Any, start: int, micro_batch_size: int) -> Any:
    def _slice(x):
        if isinstance(x, torch.Tensor):
            return x[start:start + micro_batch_size]
        return x

    return tree_map(_slice, batch)


def model_forward(model, data: Any, internal_inputs: Optional[dict]) -> Any:
    """Call the stage: user data (kwargs dict / args) + tensors produced by the previous stage."""
    internal_inputs = internal_inputs or {}
    if data is None:
        return model(**internal_inputs)
    if isinstance(data, dict):
        return model(**data, **internal_inputs)
    if isinstance(data, (tuple, list)):
        return model(*data, **internal_inputs)
    return model(data, **internal_inputs)


def retain_grad(x: Any) -> None:
    if isinstance(x, torch.Tensor) and x.requires_grad:
        x.retain_grad()


def require_grad(x: Any) -> None:
    if isinstance(x, torch.Tensor) and not x.requires_grad and x.is_floating_point():
        x.requires_grad_()


def detach(x: Any) -> Any:
    return x.detach() if isinstance(x, torch.Tensor) else x


def clone(x: Any) -> Any:
    return x.clone() if isinstance(x, torch.Tensor) else x


def release_tensor_data(x: Any) -> Any:
    if isinstance(x, torch.Tensor):
        return x.data.untyped_storage().resize_(0)
    return x


def deallocate(x: Any) -> None:
    """Free the storage of an activation that was already sent downstream (its grad_fn graph is kept alive)."""
    if isinstance(x, torch.Tensor) and x._base is None:
        x.data = torch.empty((1,), device=x.device, dtype=x.dtype)


def merge_batch(data: List[Any], batch_size_dim: int = 0) -> Any:
    if len(data) == 0:
        return None
    flattened = [tree_flatten(d) for d in data]
    spec = flattened[0][1]
    merged = []
    for leaves in zip(*[f[0] for f in flattened]):
        if isinstance(leaves[0], torch.Tensor):
            if leaves[0].dim() == 0:
                merged.append(torch.stack(leaves))
            else:
                merged.append(torch.cat(leaves, dim=batch_size_dim))
        else:
            merged.append(list(leaves))
    return tree_unflatten(merged, spec)


def default_criterion(outputs: Any, inputs: Any) -> torch.Tensor:
    """Loss of a model that computes it itself (dict with a "loss" entry, or an object with `.loss`)."""
    return outputs["loss"] if isinstance(outputs, dict) else outputs.loss
