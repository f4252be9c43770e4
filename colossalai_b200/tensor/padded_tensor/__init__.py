"""Mark a tensor as padded along one dim (vocab padding) so checkpoints store the un-padded value.
Parity: reference `colossalai/tensor/padded_tensor/api.py:56-128`."""
from __future__ import annotations

import torch

__all__ = ["is_padded_tensor", "to_padded_tensor", "to_unpadded_tensor", "init_as_padded_tensor"]


def is_padded_tensor(t: torch.Tensor) -> bool:
    return hasattr(t, "_padding_dim")


def _hijack(t: torch.Tensor) -> torch.Tensor:
    if hasattr(t, "_pad_old_detach"):
        return t
    t._pad_old_detach, t._pad_old_clone = t.detach, t.clone

    def new_detach(self=t):
        d = self._pad_old_detach()
        d._padding_dim, d._origin_length, d._current_length = self._padding_dim, self._origin_length, self._current_length
        return d

    def new_clone(self=t, *a, **k):
        c = self._pad_old_clone(*a, **k)
        c._padding_dim, c._origin_length, c._current_length = self._padding_dim, self._origin_length, self._current_length
        return c

    t.detach, t.clone = new_detach, new_clone  # type: ignore[method-assign]
    return t


def to_padded_tensor(tensor: torch.Tensor, current_length: int, padding_dim: int) -> torch.Tensor:
    if is_padded_tensor(tensor):
        return tensor
    origin = tensor.shape[padding_dim]
    pad = current_length - origin
    if pad > 0:
        shape = list(tensor.shape)
        shape[padding_dim] = pad
        tensor = torch.cat([tensor, torch.zeros(shape, dtype=tensor.dtype, device=tensor.device)], dim=padding_dim)
    tensor._padding_dim, tensor._origin_length, tensor._current_length = padding_dim, origin, current_length
    return _hijack(tensor)


def init_as_padded_tensor(tensor: torch.Tensor, current_length: int, origin_length: int, padding_dim: int):
    tensor._padding_dim, tensor._origin_length, tensor._current_length = padding_dim, origin_length, current_length
    return _hijack(tensor)


def to_unpadded_tensor(tensor: torch.Tensor) -> torch.Tensor:
    if not is_padded_tensor(tensor):
        return tensor
    sl = [slice(None)] * tensor.dim()
    sl[tensor._padding_dim] = slice(None, tensor._origin_length)
    base = tensor._pad_old_detach() if hasattr(tensor, "_pad_old_detach") else tensor
    return base[tuple(sl)].contiguous()
