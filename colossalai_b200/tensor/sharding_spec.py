"""Legacy (auto-parallel era) sharding spec: a `ShardingSpec` bound to a device mesh and a concrete tensor shape.
Parity: reference `colossalai/tensor/sharding_spec.py` (`ShardingSpec(device_mesh, entire_shape, dim_partition_dict)`,
`sharding_sequence_difference`, `get_sharded_shape_per_device`, exceptions).  Built on `tensor.d_tensor.ShardingSpec`."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .d_tensor.sharding_spec import DimSpec
from .d_tensor.sharding_spec import ShardingSpec as _Spec

__all__ = ["ShardingSpec", "ShardingSpecException", "ShardingOutOfIndexError", "DuplicatedShardingDimensionError",
           "ShardingNotDivisibleError", "_DimSpec"]

_DimSpec = DimSpec


class ShardingSpecException(Exception):
    pass


class ShardingOutOfIndexError(ShardingSpecException):
    pass


class DuplicatedShardingDimensionError(ShardingSpecException):
    pass


class ShardingNotDivisibleError(ShardingSpecException):
    pass


class ShardingSpec(_Spec):
    def __init__(self, device_mesh, entire_shape: torch.Size, dim_partition_dict: Optional[Dict[int, List[int]]] = None,
                 sharding_sequence: Optional[List[DimSpec]] = None) -> None:
        self.device_mesh = device_mesh
        self.entire_shape = torch.Size(entire_shape)
        nd = len(self.entire_shape)
        for d, axes in (dim_partition_dict or {}).items():
            if not -nd <= d < nd:
                raise ShardingOutOfIndexError(f"tensor dim {d} is out of range for a {nd}-D tensor")
            for a in axes:
                if a >= len(device_mesh.shape):
                    raise ShardingOutOfIndexError(f"mesh axis {a} does not exist in mesh {tuple(device_mesh.shape)}")
        try:
            super().__init__(nd, dim_partition_dict=dim_partition_dict, sharding_sequence=sharding_sequence)
        except AssertionError as e:
            raise DuplicatedShardingDimensionError(str(e)) from e
        for d, axes in self.dim_partition_dict.items():
            n = 1
            for a in axes:
                n *= device_mesh.shape[a]
            if self.entire_shape[d] % n != 0:
                raise ShardingNotDivisibleError(
                    f"dim {d} of size {self.entire_shape[d]} is not divisible by {n} shards")

    def sharding_sequence_difference(self, other: "ShardingSpec") -> int:
        return self.spec_diff(other)

    def get_sharded_shape_per_device(self) -> torch.Size:
        shape = list(self.entire_shape)
        for d, axes in self.dim_partition_dict.items():
            for a in axes:
                shape[d] //= self.device_mesh.shape[a]
        return torch.Size(shape)
