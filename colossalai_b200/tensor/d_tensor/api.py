"""Lightweight distributed tensor = plain tensor + attached layout / shard metadata.

Parity: reference `colossalai/tensor/d_tensor/api.py:41-531` (`shard_rowwise/colwise`, `to_global`,
`distribute_tensor`, `redistribute`, customized dtensor with user shard/gather fns for fused-QKV weights).
Two flavours: (a) mesh layouts (`Layout`) for N-D sharding; (b) the common 1-D case — a tensor sharded along one
dim over one process group — tagged with `dist_shard = (dim, group)` which is what TP layers use.
"""
from __future__ import annotations

import copy
from typing import Callable, Optional, Union

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup

from ...parallel import comm
from .layout import Layout
from .layout_converter import LayoutConverter
from .sharding_spec import ShardingSpec

_layout_converter = LayoutConverter()


# ------------------------------------------------------------------------------------------ predicates
def is_distributed_tensor(t: torch.Tensor) -> bool:
    return hasattr(t, "dist_layout") or hasattr(t, "dist_shard") or hasattr(t, "shard_fn")


def is_sharded(t: torch.Tensor) -> bool:
    if hasattr(t, "dist_shard"):
        return comm.group_size(t.dist_shard[1]) > 1
    if hasattr(t, "dist_layout"):
        return t.shape != t.dist_layout.global_shape
    return False


def is_customized_distributed_tensor(t: torch.Tensor) -> bool:
    return hasattr(t, "shard_fn") and hasattr(t, "gather_fn")


def _hijack_detach_and_clone(t: torch.Tensor) -> torch.Tensor:
    """Make `.detach()` / `.clone()` keep the distributed metadata (state_dict() detaches params)."""
    t._old_detach = t.detach
    t._old_clone = t.clone

    def new_detach(self=t):
        d = self._old_detach()
        _copy_meta(self, d)
        return d

    def new_clone(self=t, *a, **k):
        c = self._old_clone(*a, **k)
        _copy_meta(self, c)
        return c

    t.detach = new_detach  # type: ignore[method-assign]
    t.clone = new_clone  # type: ignore[method-assign]
    return t


def _copy_meta(src: torch.Tensor, dst: torch.Tensor) -> None:
    for attr in ("dist_layout", "dist_shard", "shard_fn", "gather_fn", "dist_global_shape", "tp_shard_dim"):
        if hasattr(src, attr):
            setattr(dst, attr, getattr(src, attr))


# ------------------------------------------------------------------------------------------ 1-D group sharding
def shard_along(tensor: torch.Tensor, dim: int, group: Optional[ProcessGroup]) -> torch.Tensor:
    dim = dim % tensor.dim()
    out = comm.split_along(tensor, dim, group).clone()
    out.dist_shard = (dim, group)
    out.dist_global_shape = torch.Size(tensor.shape)
    return _hijack_detach_and_clone(out)


def shard_rowwise(tensor: torch.Tensor, group_or_device_mesh=None) -> torch.Tensor:
    """Shard dim 0."""
    return shard_along(tensor, 0, _as_group(group_or_device_mesh))


def shard_colwise(tensor: torch.Tensor, group_or_device_mesh=None) -> torch.Tensor:
    """Shard the last dim."""
    return shard_along(tensor, -1, _as_group(group_or_device_mesh))


def _as_group(g):
    if g is None or isinstance(g, ProcessGroup):
        return g
    if hasattr(g, "get_group_along_axis"):
        assert len(g.shape) == 1, "shard_rowwise/colwise need a 1-D mesh or a process group"
        return g.get_group_along_axis(0)
    return g


def mark_sharded(tensor: torch.Tensor, dim: int, group: Optional[ProcessGroup], global_shape=None) -> torch.Tensor:
    """Tag an already-local shard (used when layers are constructed directly sharded / lazily)."""
    dim = dim % tensor.dim()
    tensor.dist_shard = (dim, group)
    if global_shape is None:
        gs = list(tensor.shape)
        gs[dim] *= comm.group_size(group)
        global_shape = torch.Size(gs)
    tensor.dist_global_shape = torch.Size(global_shape)
    return _hijack_detach_and_clone(tensor) if not hasattr(tensor, "_old_detach") else tensor


def sharded_tensor_to_param(t: torch.Tensor, requires_grad: bool = True) -> torch.nn.Parameter:
    p = torch.nn.Parameter(t, requires_grad=requires_grad)
    _copy_meta(t, p)
    _hijack_detach_and_clone(p)
    return p


def sharded_tensor_to_existing_param(t: torch.Tensor, param: torch.nn.Parameter) -> None:
    param.data = t
    _copy_meta(t, param)
    if not hasattr(param, "_old_detach"):
        _hijack_detach_and_clone(param)


def distribute_tensor_with_spec(global_tensor: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Shard `global_tensor` the way `like` is sharded."""
    if is_customized_distributed_tensor(like):
        return like.shard_fn(global_tensor)
    if hasattr(like, "dist_shard"):
        dim, group = like.dist_shard
        return comm.split_along(global_tensor, dim, group)
    if hasattr(like, "dist_layout"):
        return _shard_by_layout(global_tensor, like.dist_layout)
    return global_tensor


def to_global(t: torch.Tensor) -> torch.Tensor:
    """All-gather a distributed tensor back to its global value (no autograd)."""
    with torch.no_grad():
        if is_customized_distributed_tensor(t):
            return t.gather_fn(t)
        if hasattr(t, "dist_shard"):
            dim, group = t.dist_shard
            return comm.all_gather(t.detach() if not hasattr(t, "_old_detach") else t._old_detach(), dim, group)
        if hasattr(t, "dist_layout"):
            layout = t.dist_layout
            target = Layout(layout.device_mesh, ShardingSpec(layout.sharding_spec.dims, {}), layout.global_shape)
            base = t._old_detach() if hasattr(t, "_old_detach") else t.detach()
            return _layout_converter.apply(base, layout, target)
    return t


def to_global_for_customized_distributed_tensor(t: torch.Tensor) -> torch.Tensor:
    assert is_customized_distributed_tensor(t)
    return t.gather_fn(t)


# ------------------------------------------------------------------------------------------ mesh layouts
def _shard_by_layout(tensor: torch.Tensor, layout: Layout) -> torch.Tensor:
    out = tensor
    mesh = layout.device_mesh
    for dim, axes in layout.sharding_spec.dim_partition_dict.items():
        for a in axes:
            n, idx = mesh.shape[a], mesh.coordinate(a)
            out = out.chunk(n, dim=dim)[idx]
    return out.contiguous()


def distribute_tensor(tensor: torch.Tensor, device_mesh, sharding_spec: ShardingSpec) -> torch.Tensor:
    assert not is_distributed_tensor(tensor), "tensor is already distributed"
    layout = Layout(device_mesh, sharding_spec, tensor.shape)
    out = _shard_by_layout(tensor, layout).clone()
    out.dist_layout = layout
    return _hijack_detach_and_clone(out)


def init_as_dtensor(tensor: torch.Tensor, device_mesh, sharding_spec: ShardingSpec, global_shape) -> torch.Tensor:
    tensor.dist_layout = Layout(device_mesh, sharding_spec, global_shape)
    return _hijack_detach_and_clone(tensor)


def redistribute(dtensor: torch.Tensor, device_mesh, sharding_spec: ShardingSpec) -> torch.Tensor:
    assert hasattr(dtensor, "dist_layout"), "redistribute expects a mesh-layout dtensor"
    src = dtensor.dist_layout
    tgt = Layout(device_mesh, sharding_spec, src.global_shape)
    base = dtensor._old_detach() if hasattr(dtensor, "_old_detach") else dtensor
    out = _layout_converter.apply(base, src, tgt).contiguous()
    out.dist_layout = tgt
    return _hijack_detach_and_clone(out)


def get_layout(t: torch.Tensor) -> Layout:
    return t.dist_layout


def get_global_shape(t: torch.Tensor) -> torch.Size:
    if hasattr(t, "dist_global_shape"):
        return t.dist_global_shape
    if hasattr(t, "dist_layout"):
        return t.dist_layout.global_shape
    return t.shape


def get_device_mesh(t: torch.Tensor):
    return t.dist_layout.device_mesh


def get_sharding_spec(t: torch.Tensor) -> ShardingSpec:
    return t.dist_layout.sharding_spec


def is_shape_consistent(t: torch.Tensor) -> bool:
    if hasattr(t, "dist_layout"):
        return t.shape == t.dist_layout.get_sharded_shape_per_device()
    return True


# ------------------------------------------------------------------------------------------ customized dtensor
def customized_distributed_tensor_to_param(t: torch.Tensor, requires_grad: bool = True) -> torch.nn.Parameter:
    return sharded_tensor_to_param(t, requires_grad)


def distribute_tensor_with_customization(tensor: torch.Tensor, shard_fn: Callable, gather_fn: Callable) -> torch.Tensor:
    """Shard with a user function (e.g. fused-QKV: split into q/k/v blocks, shard each, re-concatenate)."""
    out = shard_fn(tensor)
    out = out.clone() if out.data_ptr() == tensor.data_ptr() else out
    out.shard_fn, out.gather_fn = shard_fn, gather_fn
    out.dist_global_shape = torch.Size(tensor.shape)
    return _hijack_detach_and_clone(out)


def mark_customized(tensor: torch.Tensor, shard_fn: Callable, gather_fn: Callable, global_shape) -> torch.Tensor:
    tensor.shard_fn, tensor.gather_fn = shard_fn, gather_fn
    tensor.dist_global_shape = torch.Size(global_shape)
    return _hijack_detach_and_clone(tensor) if not hasattr(tensor, "_old_detach") else tensor
