"""Layout = (mesh, sharding spec, global shape).  Parity: reference `colossalai/tensor/d_tensor/layout.py:12`."""
from __future__ import annotations

from typing import Tuple

import torch

from .sharding_spec import ShardingSpec


class Layout:
    def __init__(self, device_mesh, sharding_spec: ShardingSpec, global_shape: torch.Size) -> None:
        self.device_mesh = device_mesh
        self.sharding_spec = sharding_spec
        self.global_shape = torch.Size(global_shape)
        self._sanity_check()

    def _sanity_check(self) -> None:
        for dim, axes in self.sharding_spec.dim_partition_dict.items():
            n = 1
            for a in axes:
                n *= self.device_mesh.shape[a]
            assert self.global_shape[dim] % n == 0, (
                f"global dim {dim} of size {self.global_shape[dim]} is not divisible by {n} shards")

    def get_sharded_shape_per_device(self) -> torch.Size:
        shape = list(self.global_shape)
        for dim, axes in self.sharding_spec.dim_partition_dict.items():
            for a in axes:
                shape[dim] //= self.device_mesh.shape[a]
        return torch.Size(shape)

    def __repr__(self) -> str:
        return f"Layout(spec={self.sharding_spec}, global_shape={tuple(self.global_shape)}, mesh={self.device_mesh.shape})"
