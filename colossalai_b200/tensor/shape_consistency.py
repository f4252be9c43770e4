"""`ShapeConsistencyManager`: find (and price) the collective sequence that converts one sharding spec into another.
Parity: reference `colossalai/tensor/shape_consistency.py` (`shape_consistency(source, target) -> (transform_path,
comm_action_sequence, total_cost)`, `apply`, `mem_cost`, `ShapeConsistencyOptions`).  The plan comes from
`d_tensor.LayoutConverter`: a uniform-cost search over one-step transforms (gather / shard / all-to-all) priced with the
mesh's alpha-beta model - forward + backward, or forward only when `forward_only` is set - with the greedy
all-to-all -> gather -> split heuristic as the fallback."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from ..context import SingletonMeta
from .comm_spec import CommSpec
from .d_tensor.layout import Layout
from .d_tensor.layout_converter import LayoutConverter
from .sharding_spec import ShardingSpec

__all__ = ["ShapeConsistencyManager", "ShapeConsistencyOptions", "CommSpec"]


@dataclass
class ShapeConsistencyOptions:
    """`method`: "search" (cost-optimal plan) or "greedy" (deterministic heuristic)."""
    method: str = "search"


class ShapeConsistencyManager(metaclass=SingletonMeta):
    def __init__(self) -> None:
        self._options = None
        self._forward_only = False
        self.total_communication_cost = 0
        self.total_transform_steps = 0
        self.cached_spec_pairs_transform_path: Dict[Tuple[str, str], tuple] = {}
        self._converter = LayoutConverter()

    @property
    def options(self):
        return self._options

    @options.setter
    def options(self, v: ShapeConsistencyOptions) -> None:
        assert isinstance(v, ShapeConsistencyOptions)
        self._options = v

    @property
    def forward_only(self) -> bool:
        return self._forward_only

    @forward_only.setter
    def forward_only(self, v: bool) -> None:
        self._forward_only = bool(v)
        self._converter.forward_only = self._forward_only        # plans are searched under the same objective

    def mem_cost(self, transform_path: List[ShardingSpec], dtype_bytes: float = 2.0) -> float:
        """Peak bytes alive on one device while the plan runs (input + output of its widest step)."""
        if not transform_path:
            return 0.0
        mesh, shape = transform_path[0].device_mesh, transform_path[0].entire_shape
        layouts = [Layout(mesh, sp, shape) for sp in transform_path]
        return self._converter.mem_cost(layouts) * dtype_bytes

    def shape_consistency(self, source_spec: ShardingSpec, target_spec: ShardingSpec
                          ) -> Tuple[List[ShardingSpec], List[CommSpec], Dict[str, float]]:
        method = self._options.method if self._options is not None else "search"
        key = (repr(source_spec), repr(target_spec), tuple(source_spec.entire_shape), method, self._forward_only)
        if key in self.cached_spec_pairs_transform_path:
            return self.cached_spec_pairs_transform_path[key]
        mesh, shape = source_spec.device_mesh, source_spec.entire_shape
        src = Layout(mesh, source_spec, shape)
        tgt = Layout(mesh, target_spec, shape)
        layouts, comms = self._converter.layout_converting(src, tgt, method=method)
        path = [ShardingSpec(mesh, shape, dim_partition_dict=l.sharding_spec.dim_partition_dict) for l in layouts]
        cost = {"forward": 0.0, "backward": 0.0, "total": 0.0}
        nbytes_full = float(torch.Size(shape).numel()) * 2.0
        priced = []
        for spec_before, cs in zip(path[:-1], comms):
            local = nbytes_full
            for axes in spec_before.dim_partition_dict.values():
                for a in axes:
                    local /= mesh.shape[a]
            c = CommSpec(cs.comm_pattern, gather_dim=cs.gather_dim, shard_dim=cs.shard_dim,
                         logical_process_axis=cs.logical_process_axis, device_mesh=mesh)
            priced.append(c)
            step = c.get_comm_cost(local)
            for k in cost:
                cost[k] += step[k] if not (self._forward_only and k != "forward") else 0.0
        if self._forward_only:
            cost["total"] = cost["forward"]
        out = (path, priced, cost)
        self.cached_spec_pairs_transform_path[key] = out
        return out

    def apply(self, tensor_with_sharding_spec: torch.Tensor, target_spec: ShardingSpec) -> torch.Tensor:
        src: ShardingSpec = tensor_with_sharding_spec.sharding_spec
        _, comms, _ = self.shape_consistency(src, target_spec)
        t = tensor_with_sharding_spec
        for cs in comms:
            t = cs.covert_spec_to_action(t)
        t.sharding_spec = target_spec
        return t

    def apply_for_autoparallel_runtime(self, tensor: torch.Tensor, source_spec: ShardingSpec,
                                       target_spec: ShardingSpec) -> torch.Tensor:
        _, comms, _ = self.shape_consistency(source_spec, target_spec)
        for cs in comms:
            tensor = cs.covert_spec_to_action(tensor)
        return tensor
