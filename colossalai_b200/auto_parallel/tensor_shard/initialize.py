"""User entry points of the graph-level auto-parallel flow: trace -> annotate -> enumerate strategies -> ILP -> apply.

Parity: reference `colossalai/auto_parallel/tensor_shard/initialize.py:58-360` (`build_strategy_constructor`,
`solve_solution`, `transform_to_sharded_model`, `initialize_device_mesh`, `initialize_model`, `autoparallelize`).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.fx as fx
import torch.nn as nn

from ...device import DeviceMesh
from ...fx import MetaInfoProp, symbolic_trace
from .runtime import runtime_apply_pass
from .sharding_strategy import ShardingStrategy, Spec, spec_str
from .solver import CostGraph, Solver, SolverOptions, StrategiesConstructor

__all__ = ["extract_meta_args_from_dataloader", "shape_prop", "build_strategy_constructor", "solve_solution",
           "transform_to_sharded_model", "initialize_device_mesh", "initialize_model", "autoparallelize",
           "solution_summary"]


def extract_meta_args_from_dataloader(data_loader, data_process_func=None) -> Dict[str, torch.Tensor]:
    batch = next(iter(data_loader))
    if data_process_func is not None:
        batch = data_process_func(batch)
    if torch.is_tensor(batch):
        batch = {"x": batch}
    return {k: torch.empty(v.shape, dtype=v.dtype, device="meta") for k, v in batch.items() if torch.is_tensor(v)}


def shape_prop(gm: fx.GraphModule, meta_args: Dict[str, torch.Tensor]) -> None:
    """Shapes / dtypes / flops for every node without touching real memory: fake tensors stand in for the inputs and
    (through `allow_non_fake_inputs`) for the materialised parameters."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    params = list(gm.parameters())
    dev = params[0].device if params and params[0].device.type != "meta" else torch.device("cpu")
    on_meta = bool(params) and params[0].device.type == "meta"
    names = [n.target for n in gm.graph.nodes if n.op == "placeholder"]
    if on_meta:
        MetaInfoProp(gm).run(*[meta_args.get(k) for k in names])
        return
    with FakeTensorMode(allow_non_fake_inputs=True):
        args = [None if meta_args.get(k) is None else torch.empty(meta_args[k].shape, dtype=meta_args[k].dtype, device=dev)
                for k in names]
        MetaInfoProp(gm).run(*args)


def build_strategy_constructor(gm: fx.GraphModule, device_mesh: DeviceMesh,
                               solver_options: Optional[SolverOptions] = None) -> StrategiesConstructor:
    constructor = StrategiesConstructor(gm, device_mesh, solver_options)
    constructor.build_strategies_and_cost()
    return constructor


def solve_solution(gm: fx.GraphModule, constructor: StrategiesConstructor, memory_budget: float = -1.0
                   ) -> Dict[fx.Node, ShardingStrategy]:
    solver = Solver(CostGraph(constructor), memory_budget, constructor.options.time_limit_s)
    solution = solver.solve()
    gm.meta["autoparallel_objective_s"] = solver.objective
    gm.meta["autoparallel_memory_bytes"] = sum(s.memory_cost for s in solution.values())
    gm.meta["autoparallel_replicated_memory_bytes"] = sum(max(s.memory_cost for s in v)
                                                          for v in constructor.leaf_strategies)
    return solution


def transform_to_sharded_model(gm: fx.GraphModule, solution: Dict[fx.Node, ShardingStrategy],
                               device_mesh: DeviceMesh) -> Tuple[fx.GraphModule, Dict[str, Spec]]:
    return runtime_apply_pass(gm, solution, device_mesh)


def initialize_device_mesh(world_size: int = -1, physical_devices: Optional[Sequence[int]] = None,
                           logical_mesh_shape: Optional[Tuple[int, ...]] = None,
                           logical_mesh_id: Optional[torch.Tensor] = None, init_process_group: bool = True
                           ) -> DeviceMesh:
    """Default logical mesh: (1, world) on one NVSwitch domain — every factorisation prices the same there, so the
    flattest one keeps the ILP small; pass `logical_mesh_shape` for DP x TP layouts."""
    if world_size < 0:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    ids = torch.arange(world_size) if physical_devices is None else torch.tensor(list(physical_devices))
    if logical_mesh_id is not None:
        return DeviceMesh(ids, logical_mesh_id=logical_mesh_id, init_process_group=init_process_group)
    shape = tuple(logical_mesh_shape) if logical_mesh_shape is not None else (world_size,)
    return DeviceMesh(ids, shape, init_process_group=init_process_group and dist.is_initialized())


def solution_summary(gm: fx.GraphModule, solution: Dict[fx.Node, ShardingStrategy]) -> str:
    rows = [f"{'node':28s} {'strategy':24s} {'output':14s} {'compute us':>10s} {'comm us':>9s}"]
    for n in gm.graph.nodes:
        s = solution.get(n)
        if s is None:
            continue
        rows.append(f"{n.name[:28]:28s} {s.name[:24]:24s} {spec_str(s.output_spec):14s} {s.compute_cost * 1e6:10.2f} "
                    f"{s.comm_cost * 1e6:9.2f}")
    return "\n".join(rows)


def initialize_model(model: nn.Module, meta_args: Dict[str, torch.Tensor], device_mesh: DeviceMesh,
                     memory_budget: float = -1.0, solver_options: Optional[SolverOptions] = None,
                     return_solution: bool = False, apply: bool = True, leaf_modules=()):
    """Trace `model`, solve the intra-op sharding ILP on `device_mesh` and (with `apply=True`, needs the mesh's
    process groups) return the transformed GraphModule.  With `return_solution=True` also returns
    `{node name: strategy name}` and the parameter layout dictionary."""
    options = solver_options or SolverOptions()
    if memory_budget > 0:
        options.memory_budget = memory_budget
    gm = symbolic_trace(model, meta_args=meta_args, leaf_modules=leaf_modules)
    shape_prop(gm, meta_args)
    constructor = build_strategy_constructor(gm, device_mesh, options)
    solution = solve_solution(gm, constructor, options.memory_budget)
    param_specs: Dict[str, Spec] = {}
    if apply:
        gm, param_specs = transform_to_sharded_model(gm, solution, device_mesh)
    else:
        for n, s in solution.items():
            if n.op == "call_module":
                for pname, spec in s.param_specs.items():
                    if any(a is not None for a in spec):
                        param_specs[f"{n.target}.{pname}"] = tuple(spec)
    gm.meta["sharding_spec_dict"] = param_specs
    if return_solution:
        return gm, {n.name: s.name for n, s in solution.items()}, param_specs
    return gm


def autoparallelize(model: nn.Module, meta_args: Optional[Dict[str, torch.Tensor]] = None, data_loader=None,
                    data_process_func=None, logical_mesh_shape: Optional[Tuple[int, ...]] = None,
                    memory_budget: float = -1.0, solver_options: Optional[SolverOptions] = None,
                    return_solution: bool = False):
    """One call: build the mesh from the running process group, then `initialize_model`."""
    if meta_args is None:
        assert data_loader is not None, "pass meta_args or a data_loader"
        meta_args = extract_meta_args_from_dataloader(data_loader, data_process_func)
    mesh = initialize_device_mesh(logical_mesh_shape=logical_mesh_shape)
    return initialize_model(model, meta_args, mesh, memory_budget, solver_options, return_solution)
