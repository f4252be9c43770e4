"""Strategy construction, resharding costs and the ILP that picks one strategy per node.

Parity: reference `colossalai/auto_parallel/tensor_shard/solver/{strategies_constructor,cost_graph,solver}.py`
(`StrategiesConstructor`, `CostGraph`, `Solver.call_solver_serialized_args` — an Alpa-style 0/1 program solved with
PuLP).  Formulation here, solved with HiGHS through `scipy.optimize.milp`:

    minimise   sum_n sum_i x[n,i] * (compute + comm)[n,i]  +  sum_(u->v) sum_{i,j} y[uv,i,j] * reshard[uv,i,j]
    subject to sum_i x[n,i] = 1                                   one strategy per node
               y[uv,i,j] >= x[u,i] + x[v,j] - 1,  y >= 0          linearised product (costs are non-negative)
               sum_n sum_i x[n,i] * memory[n,i] <= budget         optional per-device memory budget

Only (i, j) pairs with a non-zero resharding cost get a `y` variable.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch.fx as fx

from ...device import DeviceMesh
from .node_handler import HandlerContext, generate_strategies, node_bytes, node_shape
from .sharding_strategy import ShardingStrategy, Spec, StrategiesVector, axis_dim, shard_factor

__all__ = ["SolverOptions", "StrategiesConstructor", "CostGraph", "Solver", "resharding_cost", "resharding_steps"]


@dataclass
class SolverOptions:
    """Parity: reference `tensor_shard/options.py` (`SolverOptions`: solver / dataloader / output options)."""
    shard_inputs: bool = True
    train: bool = True
    memory_budget: float = -1.0
    peak_tflops: float = 1400.0
    hbm_gbps: float = 6000.0
    time_limit_s: float = 60.0


def resharding_steps(src: Spec, dst: Spec, mesh_shape) -> List[Tuple[str, int, Optional[int], Optional[int]]]:
    """Per mesh axis the collective that turns layout `src` into `dst`: ("gather", axis, src_dim, None) /
    ("split", axis, None, dst_dim) / ("all_to_all", axis, src_dim, dst_dim).  Gathers come first so a dimension is
    free before another axis moves onto it."""
    gathers, others = [], []
    for a, n in enumerate(mesh_shape):
        if n <= 1:
            continue
        sd, dd = axis_dim(src, a), axis_dim(dst, a)
        if sd == dd:
            continue
        if dd is None:
            gathers.append(("gather", a, sd, None))
        elif sd is None:
            others.append(("split", a, None, dd))
        else:
            others.append(("all_to_all", a, sd, dd))
    return gathers + others


def resharding_cost(src: Spec, dst: Spec, nbytes: float, mesh: DeviceMesh, train: bool = True) -> float:
    """Forward + backward seconds to move a tensor of `nbytes` (global) from `src` to `dst`: a split is free forward
    and an all-gather backward, a gather the other way round, an axis changing dimension is an all-to-all each way."""
    if src == dst:
        return 0.0
    ms = tuple(mesh.shape)
    cost = 0.0
    for kind, a, sd, dd in resharding_steps(src, dst, ms):
        other = max(shard_factor(src, ms) // (ms[a] if sd is not None else 1), 1)
        part = nbytes / other
        if kind == "gather":
            cost += mesh.all_gather_cost(part, a)
        elif kind == "split":
            cost += mesh.all_gather_cost(part, a) if train else 0.0
        else:
            cost += mesh.all_to_all_cost(part, a) * (2.0 if train else 1.0)
    return cost


class StrategiesConstructor:
    """Walks the graph and attaches a `StrategiesVector` to every node (`node.meta['strategies']`)."""

    def __init__(self, gm: fx.GraphModule, mesh: DeviceMesh, options: Optional[SolverOptions] = None) -> None:
        self.gm, self.mesh = gm, mesh
        self.options = options or SolverOptions()
        self.ctx = HandlerContext(gm, mesh, self.options.peak_tflops * 1e12, self.options.hbm_gbps * 1e9,
                                  self.options.shard_inputs, self.options.train)
        self.leaf_strategies: List[StrategiesVector] = []
        self.strategy_map: Dict[fx.Node, StrategiesVector] = {}

    def build_strategies_and_cost(self) -> List[StrategiesVector]:
        for node in self.gm.graph.nodes:
            vec = generate_strategies(node, self.ctx)
            assert len(vec) > 0, f"no strategy for {node.format_node()}"
            node.meta["strategies"] = vec
            self.leaf_strategies.append(vec)
            self.strategy_map[node] = vec
        return self.leaf_strategies


class CostGraph:
    """Edge costs: `edge_costs[(u, v)][i, j]` = resharding seconds when u runs strategy i and v runs strategy j."""

    def __init__(self, constructor: StrategiesConstructor) -> None:
        self.constructor = constructor
        self.nodes: List[fx.Node] = [v.node for v in constructor.leaf_strategies]
        self.edge_costs: Dict[Tuple[fx.Node, fx.Node], np.ndarray] = {}
        mesh, train = constructor.mesh, constructor.options.train
        for v in self.nodes:
            vs: StrategiesVector = v.meta["strategies"]
            for u in v.all_input_nodes:
                if node_shape(u) is None or not any(u.name in s.input_specs for s in vs):
                    continue
                us: StrategiesVector = u.meta["strategies"]
                nbytes = node_bytes(u)
                c = np.zeros((len(us), len(vs)))
                for i, su in enumerate(us):
                    for j, sv in enumerate(vs):
                        want = sv.input_specs.get(u.name)
                        if want is None or su.output_spec is None:
                            continue
                        c[i, j] = resharding_cost(su.output_spec, want, nbytes, mesh, train)
                self.edge_costs[(u, v)] = c


class Solver:
    def __init__(self, cost_graph: CostGraph, memory_budget: float = -1.0, time_limit_s: float = 60.0) -> None:
        self.cg = cost_graph
        self.memory_budget = memory_budget
        self.time_limit_s = time_limit_s
        self.objective: Optional[float] = None

    def call_solver_serialized_args(self) -> Dict[fx.Node, int]:
        from scipy.optimize import Bounds, LinearConstraint, milp
        from scipy.sparse import lil_matrix

        nodes = self.cg.nodes
        offs, n_x = {}, 0
        for n in nodes:
            offs[n] = n_x
            n_x += len(n.meta["strategies"])
        pairs: List[Tuple[int, int, float]] = []               # (x index of u strategy, x index of v strategy, cost)
        for (u, v), c in self.cg.edge_costs.items():
            for i, j in zip(*np.nonzero(c)):
                pairs.append((offs[u] + int(i), offs[v] + int(j), float(c[i, j])))
        n_var = n_x + len(pairs)
        cost = np.zeros(n_var)
        mem = np.zeros(n_var)
        for n in nodes:
            for i, s in enumerate(n.meta["strategies"]):
                cost[offs[n] + i] = s.total_cost
                mem[offs[n] + i] = s.memory_cost
        for k, (_, _, c) in enumerate(pairs):
            cost[n_x + k] = c
        # normalise so HiGHS tolerances (absolute) do not swallow micro-second costs
        scale = 1.0 / max(cost.max(), 1e-30)
        n_rows = len(nodes) + len(pairs) + (1 if self.memory_budget > 0 else 0)
        A = lil_matrix((n_rows, n_var))
        lo, hi = np.zeros(n_rows), np.zeros(n_rows)
        r = 0
        for n in nodes:
            k = len(n.meta["strategies"])
            A[r, offs[n]: offs[n] + k] = 1.0
            lo[r] = hi[r] = 1.0
            r += 1
        for k, (xi, xj, _) in enumerate(pairs):               # x_i + x_j - y <= 1
            A[r, xi] = 1.0
            A[r, xj] = 1.0
            A[r, n_x + k] = -1.0
            lo[r], hi[r] = -np.inf, 1.0
            r += 1
        if self.memory_budget > 0:
            A[r, :n_x] = mem[:n_x]
            lo[r], hi[r] = -np.inf, self.memory_budget
            r += 1
        integrality = np.concatenate([np.ones(n_x), np.zeros(len(pairs))])
        res = milp(cost * scale, constraints=LinearConstraint(A.tocsr(), lo, hi), integrality=integrality,
                   bounds=Bounds(0.0, 1.0), options={"time_limit": self.time_limit_s, "disp": False})
        if res.x is None:
            raise RuntimeError(f"auto-parallel ILP has no feasible solution (memory budget {self.memory_budget}): "
                               f"{res.message}")
        self.objective = float(res.fun) / scale
        choice = {}
        for n in nodes:
            k = len(n.meta["strategies"])
            choice[n] = int(np.argmax(res.x[offs[n]: offs[n] + k]))
        return choice

    def solve(self) -> Dict[fx.Node, ShardingStrategy]:
        choice = self.call_solver_serialized_args()
        sol = {}
        for n, i in choice.items():
            n.meta["best_strategy"] = n.meta["strategies"][i]
            sol[n] = n.meta["best_strategy"]
        return sol
