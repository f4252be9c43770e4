"""Chain DP over per-layer sharding strategies (see package docstring)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ..device import DeviceMesh

__all__ = ["LayerStrategy", "ShardingPlan", "solve_chain", "initialize_model", "autoparallelize"]

STRATEGIES = ("replicate", "col", "row")      # col: output features sharded; row: input features sharded


@dataclass
class LayerStrategy:
    name: str                      # module path
    kind: str                      # one of STRATEGIES
    compute_s: float
    comm_s: float
    param_bytes: float
    out_sharded: bool              # activation leaving the layer is feature-sharded
    in_sharded: bool               # layer expects a feature-sharded input
    out_features: int = 0          # width of the activation leaving the layer


@dataclass
class ShardingPlan:
    strategies: List[LayerStrategy] = field(default_factory=list)
    total_time_s: float = 0.0
    param_bytes_per_device: float = 0.0

    def as_dict(self) -> Dict[str, str]:
        return {s.name: s.kind for s in self.strategies}


def _layer_candidates(name: str, m: nn.Module, tokens: int, mesh: DeviceMesh, axis: int, peak_flops: float,
                      elem: int) -> List[LayerStrategy]:
    n = mesh.shape[axis]
    if isinstance(m, nn.Linear):
        fin, fout = m.in_features, m.out_features
        flops = 2.0 * tokens * fin * fout * 3          # fwd + dgrad + wgrad
        pbytes = fin * fout * elem
        act_out = tokens * fout * elem
        act_in = tokens * fin * elem
        return [
            LayerStrategy(name, "replicate", flops / peak_flops, 0.0, pbytes, False, False, fout),
            # column: no fwd comm, dgrad needs an all-reduce of the input gradient
            LayerStrategy(name, "col", flops / n / peak_flops, mesh.all_reduce_cost(act_in, axis), pbytes / n, True, False,
                          fout),
            # row: fwd all-reduce of the output
            LayerStrategy(name, "row", flops / n / peak_flops, mesh.all_reduce_cost(act_out, axis), pbytes / n, False, True,
                          fout),
        ]
    if isinstance(m, nn.Embedding):
        pbytes = m.num_embeddings * m.embedding_dim * elem
        act = tokens * m.embedding_dim * elem
        return [LayerStrategy(name, "replicate", 0.0, 0.0, pbytes, False, False, m.embedding_dim),
                LayerStrategy(name, "row", 0.0, mesh.all_reduce_cost(act, axis), pbytes / n, False, False,
                              m.embedding_dim)]
    return []


def solve_chain(layers: List[List[LayerStrategy]], mesh: DeviceMesh, axis: int, tokens: int, elem: int = 2,
                memory_budget: float = -1.0) -> ShardingPlan:
    """Exact DP over a chain: cost = compute + comm + resharding between neighbours (gather when a feature-sharded
    output meets a layer that wants the full features; nothing when col -> row pairs up)."""
    INF = float("inf")
    if not layers:
        return ShardingPlan()

    def reshard(prev: LayerStrategy, cur: LayerStrategy, width_bytes: float) -> float:
        if prev.out_sharded and not cur.in_sharded:
            return mesh.all_gather_cost(width_bytes, axis)
        if not prev.out_sharded and cur.in_sharded:
            return 0.0          # local slice
        return 0.0

    # Lagrangian sweep on the memory budget: add lambda * param_bytes to the cost until the plan fits
    def run(lmbda: float) -> Tuple[float, List[int]]:
        best = [[INF] * len(c) for c in layers]
        back = [[-1] * len(c) for c in layers]
        for j, s in enumerate(layers[0]):
            best[0][j] = s.compute_s + s.comm_s + lmbda * s.param_bytes
        for i in range(1, len(layers)):
            for j, s in enumerate(layers[i]):
                for k, p in enumerate(layers[i - 1]):
                    c = best[i - 1][k] + reshard(p, s, float(tokens) * elem * max(p.out_features, 1)) + s.compute_s + s.comm_s \
                        + lmbda * s.param_bytes
                    if c < best[i][j]:
                        best[i][j], back[i][j] = c, k
        j = min(range(len(layers[-1])), key=lambda q: best[-1][q])
        choice = [j]
        for i in range(len(layers) - 1, 0, -1):
            j = back[i][j]
            choice.append(j)
        return best[-1][choice[0]], choice[::-1]

    lmbda, plan_choice = 0.0, None
    for _ in range(40):
        _, choice = run(lmbda)
        mem = sum(layers[i][j].param_bytes for i, j in enumerate(choice))
        plan_choice = choice
        if memory_budget <= 0 or mem <= memory_budget:
            break
        lmbda = max(lmbda * 2, 1e-12)
    strategies = [layers[i][j] for i, j in enumerate(plan_choice)]
    total = sum(s.compute_s + s.comm_s for s in strategies)
    return ShardingPlan(strategies, total, sum(s.param_bytes for s in strategies))


def initialize_model(model: nn.Module, meta_args: Dict[str, torch.Tensor], device_mesh: DeviceMesh,
                     memory_budget: float = -1.0, mesh_axis: int = -1, peak_tflops: float = 1400.0,
                     return_solution: bool = True) -> ShardingPlan:
    """Pick a sharding strategy for every Linear / Embedding of `model` (a chain in registration order)."""
    axis = mesh_axis % len(device_mesh.shape)
    ids = next(iter(meta_args.values()))
    tokens = int(ids.numel()) if ids.dim() <= 2 else int(ids.shape[0] * ids.shape[1])
    elem = 2
    layers = []
    for name, m in model.named_modules():
        cands = _layer_candidates(name, m, tokens, device_mesh, axis, peak_tflops * 1e12, elem)
        if cands:
            layers.append(cands)
    return solve_chain(layers, device_mesh, axis, tokens, elem, memory_budget)


def autoparallelize(model: nn.Module, meta_args: Dict[str, torch.Tensor], device_mesh: DeviceMesh, process_group=None,
                    memory_budget: float = -1.0, **kw) -> Tuple[nn.Module, ShardingPlan]:
    """Solve, then apply the plan with the Shardformer 1-D parallel linears (needs an initialised process group for the
    chosen mesh axis; with `process_group=None` only the plan is returned and the model is left untouched)."""
    plan = initialize_model(model, meta_args, device_mesh, memory_budget, **kw)
    if process_group is None:
        return model, plan
    from ..shardformer.layer import Linear1D_Col, Linear1D_Row

    mods = dict(model.named_modules())
    for s in plan.strategies:
        m = mods[s.name]
        if not isinstance(m, nn.Linear) or s.kind == "replicate":
            continue
        parent_name, _, child = s.name.rpartition(".")
        parent = mods[parent_name] if parent_name else model
        cls = Linear1D_Col if s.kind == "col" else Linear1D_Row
        kwargs = dict(gather_output=True) if s.kind == "col" else dict(parallel_input=False)
        setattr(parent, child, cls.from_native_module(m, process_group, **kwargs))
    return model, plan
