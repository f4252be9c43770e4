"""Sharding specs and per-node strategies of the graph-level intra-op solver.

Parity: reference `colossalai/auto_parallel/tensor_shard/sharding_strategy.py` (`ShardingStrategy`,
`StrategiesVector`, `TrainCycleItem`, `MemoryCost`).  A spec here is a plain tuple with one entry per tensor dimension:
the logical mesh axis that dimension is split over, or `None`.  A mesh axis appears at most once in a spec; an axis
that does not appear means "replicated along that axis".  Partial sums never travel along an edge: a strategy that
produces one (row-parallel linear) owns the all-reduce and reports it in its communication cost.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch.fx as fx

__all__ = ["Spec", "ShardingStrategy", "StrategiesVector", "replicated", "enumerate_specs", "spec_str", "shard_factor",
           "local_shape", "axis_dim"]

Spec = Tuple[Optional[int], ...]


def replicated(ndim: int) -> Spec:
    return (None,) * ndim


def spec_str(spec: Optional[Spec]) -> str:
    if spec is None:
        return "-"
    return "[" + ",".join("R" if a is None else f"S{a}" for a in spec) + "]" if spec else "[]"


def axis_dim(spec: Spec, axis: int) -> Optional[int]:
    """Tensor dimension that is split over mesh `axis` (None: replicated along it)."""
    for d, a in enumerate(spec):
        if a == axis:
            return d
    return None


def shard_factor(spec: Optional[Spec], mesh_shape: Sequence[int]) -> int:
    f = 1
    for a in spec or ():
        if a is not None:
            f *= mesh_shape[a]
    return f


def local_shape(shape: Sequence[int], spec: Spec, mesh_shape: Sequence[int]) -> Tuple[int, ...]:
    return tuple(s if a is None else s // mesh_shape[a] for s, a in zip(shape, spec))


def enumerate_specs(shape: Sequence[int], mesh_shape: Sequence[int], allowed_dims: Optional[Sequence[int]] = None
                    ) -> List[Spec]:
    """Every way of laying `shape` over the mesh: each mesh axis (of size > 1) goes to one divisible dimension of
    `allowed_dims` or stays unused; one tensor dimension carries at most one axis."""
    nd = len(shape)
    dims = list(range(nd)) if allowed_dims is None else [d % nd for d in allowed_dims]
    axes = [a for a, n in enumerate(mesh_shape) if n > 1]
    out: List[Spec] = []
    for assign in itertools.product([None] + dims, repeat=len(axes)):
        used = [d for d in assign if d is not None]
        if len(set(used)) != len(used):
            continue
        spec: List[Optional[int]] = [None] * nd
        ok = True
        for a, d in zip(axes, assign):
            if d is None:
                continue
            if shape[d] % mesh_shape[a] or shape[d] < mesh_shape[a]:
                ok = False
                break
            spec[d] = a
        if ok:
            out.append(tuple(spec))
    return out


@dataclass
class ShardingStrategy:
    """One way of executing a node on the mesh.

    `input_specs` maps a producer node name to the layout this strategy wants for that operand.
    `reduce_axes`: the node's raw result is a partial sum over these mesh axes (all-reduced in forward by the runtime).
    `bwd_reduce_inputs`: operand name -> mesh axes along which the operand is replicated while the computation is
    not, i.e. its gradient arrives as a partial sum and is all-reduced in backward.
    `grad_sync_axes`: same statement for the node's own parameters (handled by gradient hooks).
    `param_specs`: layout of the node's parameters (`weight`, `bias`).
    """
    name: str
    output_spec: Optional[Spec]
    input_specs: Dict[str, Spec] = field(default_factory=dict)
    param_specs: Dict[str, Spec] = field(default_factory=dict)
    compute_cost: float = 0.0
    comm_cost: float = 0.0
    memory_cost: float = 0.0
    reduce_axes: Tuple[int, ...] = ()
    bwd_reduce_inputs: Dict[str, Tuple[int, ...]] = field(default_factory=dict)
    grad_sync_axes: Tuple[int, ...] = ()

    @property
    def total_cost(self) -> float:
        return self.compute_cost + self.comm_cost

    def __repr__(self) -> str:
        ins = ", ".join(f"{k}:{spec_str(v)}" for k, v in self.input_specs.items())
        return (f"Strategy({self.name}: ({ins}) -> {spec_str(self.output_spec)}, compute {self.compute_cost:.2e}s, "
                f"comm {self.comm_cost:.2e}s, mem {self.memory_cost:.3g}B)")


class StrategiesVector(list):
    """All candidate strategies of one fx node."""

    def __init__(self, node: fx.Node) -> None:
        super().__init__()
        self.node = node

    def names(self) -> List[str]:
        return [s.name for s in self]
