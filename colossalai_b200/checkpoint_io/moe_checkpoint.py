"""Checkpoint IO for expert-parallel models.

Parity: reference `colossalai/checkpoint_io/moe_checkpoint.py:44-860` (expert params gathered over the ep group before
the (dp 0, tp 0) writer stores them; optimizer states of expert params gathered over moe_dp (ZeRO) then ep).

Here expert parameters carry the same `dist_shard=(0, ep_group)` tag as tensor-parallel shards, so the generic
gather (`to_global`) / re-shard (`distribute_tensor_with_spec`) code of `HybridParallelCheckpointIO` handles them; this
class only records the extra groups and keeps the reference's entry points.
"""
from __future__ import annotations

from typing import Optional

import torch.distributed as dist
from torch.distributed import ProcessGroup

from .hybrid_parallel_checkpoint_io import HybridParallelCheckpointIO

__all__ = ["MoECheckpointIO"]


class MoECheckpointIO(HybridParallelCheckpointIO):
    def __init__(self, global_dp_group: ProcessGroup, pp_group: ProcessGroup, tp_group: ProcessGroup,
                 sp_group: ProcessGroup, ep_group: ProcessGroup, moe_dp_group: ProcessGroup, zero_stage: int,
                 verbose: bool = True) -> None:
        super().__init__(global_dp_group, pp_group, tp_group, sp_group, zero_stage, verbose)
        self.ep_group, self.moe_dp_group = ep_group, moe_dp_group
        self.ep_size = dist.get_world_size(ep_group) if ep_group is not None else 1
        self.ep_rank = dist.get_rank(ep_group) if ep_group is not None else 0
        self.moe_dp_size = dist.get_world_size(moe_dp_group) if moe_dp_group is not None else 1
        self.moe_dp_rank = dist.get_rank(moe_dp_group) if moe_dp_group is not None else 0

    def pre_save_model(self, model) -> dict:
        """Full (ep-gathered) state dict on every rank — the reference's helper for unsharded saves."""
        module_of = self._module_of(model)
        from .hybrid_parallel_checkpoint_io import _unpadded_global

        return {n: _unpadded_global(p, module_of).cpu() for n, p in model.named_parameters()}
