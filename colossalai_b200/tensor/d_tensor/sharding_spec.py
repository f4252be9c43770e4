"""Sharding spec: which mesh axes partition which tensor dim.
Parity: reference `colossalai/tensor/d_tensor/sharding_spec.py:14,156` (DimSpec / ShardingSpec, "R", "S0", "S01")."""
from __future__ import annotations

from typing import Dict, List, Optional


class DimSpec:
    """Sharding of ONE tensor dim: replicated ("R") or split over mesh axes ("S0", "S1", "S01")."""

    def __init__(self, shard_list: Optional[List[int]] = None) -> None:
        self.shard_list = sorted(shard_list or [])
        self.is_replica = len(self.shard_list) == 0

    def __eq__(self, other) -> bool:
        return isinstance(other, DimSpec) and self.shard_list == other.shard_list

    def __hash__(self):
        return hash(tuple(self.shard_list))

    def __repr__(self) -> str:
        return "R" if self.is_replica else "S" + "".join(str(a) for a in self.shard_list)


class ShardingSpec:
    def __init__(self, dim_size: int, dim_partition_dict: Optional[Dict[int, List[int]]] = None,
                 sharding_sequence: Optional[List[DimSpec]] = None) -> None:
        self.dims = dim_size
        if sharding_sequence is not None:
            assert len(sharding_sequence) == dim_size
            self.sharding_sequence = list(sharding_sequence)
            self.dim_partition_dict = {i: list(s.shard_list) for i, s in enumerate(sharding_sequence) if not s.is_replica}
        else:
            self.dim_partition_dict = {int(k) % dim_size: list(v) for k, v in (dim_partition_dict or {}).items() if v}
            self.sharding_sequence = [DimSpec(self.dim_partition_dict.get(i)) for i in range(dim_size)]
        used = [a for v in self.dim_partition_dict.values() for a in v]
        assert len(used) == len(set(used)), f"a mesh axis may shard only one tensor dim: {self.dim_partition_dict}"

    def __eq__(self, other) -> bool:
        return isinstance(other, ShardingSpec) and self.sharding_sequence == other.sharding_sequence

    def __repr__(self) -> str:
        return "[" + ", ".join(repr(s) for s in self.sharding_sequence) + "]"

    def spec_diff(self, other: "ShardingSpec") -> int:
        """Number of tensor dims whose sharding differs (a crude conversion distance)."""
        return sum(1 for a, b in zip(self.sharding_sequence, other.sharding_sequence) if a != b)
