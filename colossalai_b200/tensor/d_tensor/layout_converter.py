"""Convert a tensor between two layouts with a sequence of gather / split / all-to-all steps.

Parity: reference `colossalai/tensor/d_tensor/layout_converter.py:39` (greedy search over one-step transforms).
Here the search is a simple deterministic heuristic that is optimal on an NVSwitch box (uniform bandwidth):
  1. where source shards dim i on axis a and target shards dim j!=i on the same axis a -> one all-to-all;
  2. remaining source shards not in target -> all-gather;  3. remaining target shards -> local split.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .comm_spec import CollectiveCommPattern, CommSpec
from .layout import Layout
from .sharding_spec import ShardingSpec


class LayoutConverter:
    def __init__(self) -> None:
        self.cached_solution = {}

    def layout_converting(self, source: Layout, target: Layout) -> Tuple[List[Layout], List[CommSpec]]:
        assert source.global_shape == target.global_shape
        mesh = source.device_mesh
        key = (repr(source.sharding_spec), repr(target.sharding_spec), tuple(source.global_shape))
        if key in self.cached_solution:
            return self.cached_solution[key]
        cur = {d: list(a) for d, a in source.sharding_spec.dim_partition_dict.items()}
        tgt = {d: list(a) for d, a in target.sharding_spec.dim_partition_dict.items()}
        nd = source.sharding_spec.dims
        path, comms = [source], []

        def emit(spec_dict, cs):
            comms.append(cs)
            path.append(Layout(mesh, ShardingSpec(nd, dim_partition_dict=spec_dict), source.global_shape))

        # 1) all-to-all moves
        for d_src in list(cur.keys()):
            for axis in list(cur.get(d_src, [])):
                d_tgt = next((d for d, axes in tgt.items() if axis in axes), None)
                if d_tgt is not None and d_tgt != d_src and cur[d_src][-1] == axis:
                    cur[d_src].remove(axis)
                    if not cur[d_src]:
                        del cur[d_src]
                    cur.setdefault(d_tgt, []).append(axis)
                    emit({k: list(v) for k, v in cur.items()},
                         CommSpec(CollectiveCommPattern.ALL2ALL_FWD_ALL2ALL_BWD, gather_dim=d_src, shard_dim=d_tgt,
                                  logical_process_axis=axis, device_mesh=mesh))
        # 2) gathers (innermost axis first)
        for d in list(cur.keys()):
            while d in cur and cur[d] and cur[d] != tgt.get(d, [])[: len(cur[d])]:
                axis = cur[d].pop()
                if not cur[d]:
                    del cur[d]
                emit({k: list(v) for k, v in cur.items()},
                     CommSpec(CollectiveCommPattern.GATHER_FWD_SPLIT_BWD, gather_dim=d, logical_process_axis=axis,
                              device_mesh=mesh))
        # 3) splits
        for d, axes in tgt.items():
            have = cur.get(d, [])
            for axis in axes[len(have):]:
                cur.setdefault(d, []).append(axis)
                emit({k: list(v) for k, v in cur.items()},
                     CommSpec(CollectiveCommPattern.SPLIT_FWD_GATHER_BWD, shard_dim=d, logical_process_axis=axis,
                              device_mesh=mesh))
        assert ShardingSpec(nd, dim_partition_dict=cur) == target.sharding_spec, (cur, tgt)
        self.cached_solution[key] = (path, comms)
        return path, comms

    def apply(self, tensor: torch.Tensor, source: Layout, target: Layout) -> torch.Tensor:
        _, comms = self.layout_converting(source, target)
        for cs in comms:
            tensor = cs.covert_spec_to_action(tensor)
        return tensor
