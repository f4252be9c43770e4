"""Communication actions between layouts.  Parity: reference `colossalai/tensor/d_tensor/comm_spec.py:14-302`
(CollectiveCommPattern GATHER / ALL2ALL / SPLIT / ALLREDUCE + `covert_spec_to_action`)."""
from __future__ import annotations

from enum import Enum

import torch

from ...parallel import comm


class CollectiveCommPattern(Enum):
    GATHER_FWD_SPLIT_BWD = "gather_fwd_split_bwd"
    ALL2ALL_FWD_ALL2ALL_BWD = "all2all_fwd_all2all_bwd"
    SPLIT_FWD_GATHER_BWD = "split_fwd_gather_bwd"
    ALLREDUCE_FWD_IDENTITY_BWD = "all_reduce_fwd_identity_bwd"
    IDENTITY_FWD_ALLREDUCE_BWD = "identity_fwd_all_reduce_bwd"


class CommSpec:
    """One collective step applied along `logical_process_axis` of the mesh."""

    def __init__(self, comm_pattern: CollectiveCommPattern, process_group_dict=None, gather_dim: int = None,
                 shard_dim: int = None, logical_process_axis: int = None, device_mesh=None) -> None:
        self.comm_pattern = comm_pattern
        self.gather_dim, self.shard_dim = gather_dim, shard_dim
        self.logical_process_axis = logical_process_axis
        self.device_mesh = device_mesh
        self.process_group_dict = process_group_dict

    def _group(self):
        if self.process_group_dict is not None:
            return self.process_group_dict[self.logical_process_axis]
        return self.device_mesh.get_group_along_axis(self.logical_process_axis)

    def __repr__(self) -> str:
        return (f"CommSpec({self.comm_pattern.value}, gather_dim={self.gather_dim}, shard_dim={self.shard_dim}, "
                f"axis={self.logical_process_axis})")

    def covert_spec_to_action(self, tensor: torch.Tensor) -> torch.Tensor:
        from ...shardformer.layer import _operation as op

        g = self._group()
        p = self.comm_pattern
        if p == CollectiveCommPattern.GATHER_FWD_SPLIT_BWD:
            return op.gather_forward_split_backward(tensor, self.gather_dim, g)
        if p == CollectiveCommPattern.SPLIT_FWD_GATHER_BWD:
            return op.split_forward_gather_backward(tensor, self.shard_dim, g)
        if p == CollectiveCommPattern.ALL2ALL_FWD_ALL2ALL_BWD:
            return op.all_to_all_comm(tensor, g, scatter_dim=self.shard_dim, gather_dim=self.gather_dim)
        if p == CollectiveCommPattern.ALLREDUCE_FWD_IDENTITY_BWD:
            return op.reduce_forward(tensor, g)
        if p == CollectiveCommPattern.IDENTITY_FWD_ALLREDUCE_BWD:
            return op.reduce_backward(tensor, g)
        raise ValueError(p)
