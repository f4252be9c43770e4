"""Placement of a tensor over the tensor-parallel group (reference `legacy/tensor/distspec.py`)."""
from enum import Enum
from typing import List

__all__ = ["DistPlacementPattern", "ReplicaSpec", "ShardSpec"]


class DistPlacementPattern(Enum):
    REPLICATE = "r"
    SHARD = "s"


class _DistSpec:
    def __init__(self, placement: DistPlacementPattern, **meta_info) -> None:
        self.placement = placement
        for k, v in meta_info.items():
            setattr(self, k, v)
        self._keys = tuple(meta_info)

    def __eq__(self, other) -> bool:
        if not isinstance(other, _DistSpec) or self.placement != other.placement:
            return False
        return all(getattr(self, k) == getattr(other, k, None) for k in self._keys) and self._keys == other._keys

    def __hash__(self) -> int:
        return hash((self.placement, tuple(getattr(self, k) for k in self._keys)))

    def __repr__(self) -> str:
        body = ", ".join(f"{k}={getattr(self, k)}" for k in self._keys)
        return f"DistSpec({self.placement.name}{', ' + body if body else ''})"


def ReplicaSpec() -> _DistSpec:
    return _DistSpec(DistPlacementPattern.REPLICATE)


def ShardSpec(dims: List[int], num_partitions: List[int]) -> _DistSpec:
    """Dimension `dims[i]` is cut into `num_partitions[i]` pieces; the product must equal the TP group size."""
    assert isinstance(dims, (list, tuple)) and isinstance(num_partitions, (list, tuple)) and len(dims) == len(num_partitions)
    return _DistSpec(DistPlacementPattern.SHARD, dims=tuple(dims), num_partitions=tuple(num_partitions))
