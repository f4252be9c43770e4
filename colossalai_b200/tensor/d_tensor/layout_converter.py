"""Convert a tensor between two layouts with a sequence of gather / split / all-to-all steps.

Parity: reference `colossalai/tensor/d_tensor/layout_converter.py:39-620` (search over one-step transforms with a
heuristic, cached solutions, `layout_converting`, `apply`).

Two planners:
  * `method="search"` (default): uniform-cost search (Dijkstra) over sharding states.  One-step transforms of a state
        gather   - drop the LAST mesh axis of a sharded tensor dim (all-gather along it);
        shard    - split a tensor dim along a FREE mesh axis (local slice; an all-gather in backward);
        all2all  - move the last mesh axis of one tensor dim to the end of another dim's axis list;
    are priced with the mesh's alpha-beta model (`all_gather_cost` / `all_to_all_cost`, bytes = the local shard that
    enters the collective; forward + backward unless `forward_only`), so the plan is the cheapest sequence, with fewer
    steps winning ties.  Axis lists are kept in increasing order (the canonical "S01" form of `DimSpec`): a transform
    that would create a physically different nesting than the tag says is not generated.
  * `method="greedy"`: the deterministic all-to-all -> gather -> split heuristic (optimal on a uniform NVSwitch mesh for
    the common TP / SP conversions; kept as the fallback for meshes without a cost model and for very large state
    spaces).
`mem_cost(path)` gives the peak number of local elements alive during a plan (input + output of the widest step).
"""
from __future__ import annotations

import heapq
import itertools
from typing import Dict, List, Optional, Tuple

import torch

from .comm_spec import CollectiveCommPattern, CommSpec
from .layout import Layout
from .sharding_spec import ShardingSpec

_State = Tuple[Tuple[int, ...], ...]          # per tensor dim: the mesh axes that shard it, in nesting order


def _state_of(spec: ShardingSpec) -> _State:
    return tuple(tuple(spec.dim_partition_dict.get(d, ())) for d in range(spec.dims))


def _dict_of(state: _State) -> Dict[int, List[int]]:
    return {d: list(a) for d, a in enumerate(state) if a}


class LayoutConverter:
    def __init__(self, forward_only: bool = False, method: str = "search") -> None:
        self.cached_solution = {}
        self.forward_only = forward_only
        self.method = method

    # ------------------------------------------------------------------ cost model
    @staticmethod
    def _mesh_sizes(mesh) -> Tuple[int, ...]:
        shape = getattr(mesh, "shape", None)
        if shape is None:
            shape = getattr(mesh, "_mesh_shape", ())
        return tuple(int(s) for s in shape)

    def _local_elems(self, state: _State, global_shape, sizes) -> float:
        n = float(torch.Size(global_shape).numel())
        for axes in state:
            for a in axes:
                n /= sizes[a]
        return n

    def _step_cost(self, mesh, kind: str, axis: int, elems_in: float, elems_out: float, elem_bytes: float = 2.0) -> float:
        """Seconds (alpha-beta model of an analytical mesh) or a unit-less proxy with the same ordering."""
        has_model = hasattr(mesh, "all_gather_cost") and hasattr(mesh, "all_to_all_cost")
        n = self._mesh_sizes(mesh)[axis]

        def gather(nbytes_out):
            return mesh.all_gather_cost(nbytes_out, axis) if has_model else 1.0 + nbytes_out * (n - 1) / n * 1e-9

        def a2a(nbytes):
            return mesh.all_to_all_cost(nbytes, axis) if has_model else 1.0 + nbytes * (n - 1) / n / 2 * 1e-9

        if kind == "gather":                       # forward all-gather; backward is a local slice
            return gather(elems_out * elem_bytes)
        if kind == "shard":                        # forward local slice; backward all-gather of the gradient
            return 0.0 if self.forward_only else gather(elems_in * elem_bytes)
        cost = a2a(elems_in * elem_bytes)          # all-to-all both ways
        return cost if self.forward_only else 2 * cost

    # ------------------------------------------------------------------ one-step transforms
    def _neighbours(self, state: _State, global_shape, sizes):
        used = {a for axes in state for a in axes}
        free = [a for a in range(len(sizes)) if a not in used]
        nd = len(state)
        for d in range(nd):
            axes = state[d]
            if axes:                               # gather the innermost axis of dim d
                a = axes[-1]
                new = list(state)
                new[d] = axes[:-1]
                yield tuple(new), ("gather", d, None, a)
                for t in range(nd):                # all-to-all: axis a moves from dim d to dim t
                    if t == d:
                        continue
                    taxes = state[t]
                    if taxes and taxes[-1] > a:
                        continue                   # keep the canonical increasing nesting order
                    local_t = global_shape[t]
                    for x in taxes:
                        local_t //= sizes[x]
                    if local_t % sizes[a] != 0:
                        continue
                    new = list(state)
                    new[d] = axes[:-1]
                    new[t] = taxes + (a,)
                    yield tuple(new), ("all2all", d, t, a)
            for a in free:                         # shard dim d along a free axis
                if axes and axes[-1] > a:
                    continue
                local_d = global_shape[d]
                for x in axes:
                    local_d //= sizes[x]
                if local_d % sizes[a] != 0:
                    continue
                new = list(state)
                new[d] = axes + (a,)
                yield tuple(new), ("shard", None, d, a)

    # ------------------------------------------------------------------ planners
    def _search(self, source: Layout, target: Layout) -> Optional[List[Tuple[_State, tuple]]]:
        mesh, shape = source.device_mesh, tuple(source.global_shape)
        sizes = self._mesh_sizes(mesh)
        if not sizes:
            return None
        start, goal = _state_of(source.sharding_spec), _state_of(target.sharding_spec)
        if any(list(a) != sorted(a) for a in itertools.chain(start, goal)):
            return None                            # non-canonical nesting: leave it to the greedy planner
        counter = itertools.count()
        heap = [(0.0, 0, next(counter), start, [])]
        best: Dict[_State, Tuple[float, int]] = {start: (0.0, 0)}
        expanded = 0
        while heap:
            cost, steps, _, state, trail = heapq.heappop(heap)
            if state == goal:
                return trail
            if best.get(state, (float("inf"), 0)) < (cost, steps):
                continue
            expanded += 1
            if expanded > 20000:
                return None
            e_in = self._local_elems(state, shape, sizes)
            for nxt, action in self._neighbours(state, shape, sizes):
                e_out = self._local_elems(nxt, shape, sizes)
                c = cost + self._step_cost(mesh, action[0], action[3], e_in, e_out)
                key = (c, steps + 1)
                if key < best.get(nxt, (float("inf"), 0)):
                    best[nxt] = key
                    heapq.heappush(heap, (c, steps + 1, next(counter), nxt, trail + [(nxt, action)]))
        return None

    def _greedy(self, source: Layout, target: Layout) -> List[Tuple[_State, tuple]]:
        cur = {d: list(a) for d, a in source.sharding_spec.dim_partition_dict.items()}
        tgt = {d: list(a) for d, a in target.sharding_spec.dim_partition_dict.items()}
        nd = source.sharding_spec.dims
        trail: List[Tuple[_State, tuple]] = []

        def snap() -> _State:
            return tuple(tuple(cur.get(d, ())) for d in range(nd))

        # 1) all-to-all moves
        for d_src in list(cur.keys()):
            for axis in list(cur.get(d_src, [])):
                d_tgt = next((d for d, axes in tgt.items() if axis in axes), None)
                # the axis may only be appended where the target nests it: directly after the axes the target dim
                # already holds in target order (otherwise the tag and the physical nesting would disagree)
                fits = d_tgt is not None and tgt[d_tgt][: len(cur.get(d_tgt, [])) + 1] == cur.get(d_tgt, []) + [axis]
                if d_tgt is not None and d_tgt != d_src and cur[d_src][-1] == axis and fits:
                    cur[d_src].remove(axis)
                    if not cur[d_src]:
                        del cur[d_src]
                    cur.setdefault(d_tgt, []).append(axis)
                    trail.append((snap(), ("all2all", d_src, d_tgt, axis)))
        # 2) gathers (innermost axis first)
        for d in list(cur.keys()):
            while d in cur and cur[d] and cur[d] != tgt.get(d, [])[: len(cur[d])]:
                axis = cur[d].pop()
                if not cur[d]:
                    del cur[d]
                trail.append((snap(), ("gather", d, None, axis)))
        # 3) splits
        for d, axes in tgt.items():
            have = cur.get(d, [])
            for axis in axes[len(have):]:
                cur.setdefault(d, []).append(axis)
                trail.append((snap(), ("shard", None, d, axis)))
        assert {d: a for d, a in cur.items() if a} == {d: a for d, a in tgt.items() if a}, (cur, tgt)
        return trail

    def layout_converting(self, source: Layout, target: Layout, method: Optional[str] = None
                          ) -> Tuple[List[Layout], List[CommSpec]]:
        assert source.global_shape == target.global_shape
        mesh = source.device_mesh
        method = method or self.method
        key = (repr(source.sharding_spec), repr(target.sharding_spec), tuple(source.global_shape), method,
               self.forward_only, id(mesh))
        if key in self.cached_solution:
            return self.cached_solution[key]
        trail = self._search(source, target) if method == "search" else None
        if trail is None:
            trail = self._greedy(source, target)
        nd = source.sharding_spec.dims
        path, comms = [source], []
        for state, (kind, d_from, d_to, axis) in trail:
            if kind == "gather":
                cs = CommSpec(CollectiveCommPattern.GATHER_FWD_SPLIT_BWD, gather_dim=d_from, logical_process_axis=axis,
                              device_mesh=mesh)
            elif kind == "shard":
                cs = CommSpec(CollectiveCommPattern.SPLIT_FWD_GATHER_BWD, shard_dim=d_to, logical_process_axis=axis,
                              device_mesh=mesh)
            else:
                cs = CommSpec(CollectiveCommPattern.ALL2ALL_FWD_ALL2ALL_BWD, gather_dim=d_from, shard_dim=d_to,
                              logical_process_axis=axis, device_mesh=mesh)
            comms.append(cs)
            path.append(Layout(mesh, ShardingSpec(nd, dim_partition_dict=_dict_of(state)), source.global_shape))
        assert path[-1].sharding_spec == target.sharding_spec
        self.cached_solution[key] = (path, comms)
        return path, comms

    def plan_cost(self, source: Layout, target: Layout, method: Optional[str] = None) -> float:
        """Total modelled cost of the plan `layout_converting` returns (same cost model as the search)."""
        path, comms = self.layout_converting(source, target, method)
        sizes = self._mesh_sizes(source.device_mesh)
        total = 0.0
        for before, after, cs in zip(path[:-1], path[1:], comms):
            kind = {CollectiveCommPattern.GATHER_FWD_SPLIT_BWD: "gather", CollectiveCommPattern.SPLIT_FWD_GATHER_BWD: "shard",
                    CollectiveCommPattern.ALL2ALL_FWD_ALL2ALL_BWD: "all2all"}[cs.comm_pattern]
            total += self._step_cost(source.device_mesh, kind, cs.logical_process_axis,
                                     self._local_elems(_state_of(before.sharding_spec), source.global_shape, sizes),
                                     self._local_elems(_state_of(after.sharding_spec), source.global_shape, sizes))
        return total

    def mem_cost(self, path: List[Layout]) -> float:
        """Peak number of local elements alive while the plan runs: input + output of its widest step."""
        if len(path) < 2:
            return self._local_elems(_state_of(path[0].sharding_spec), path[0].global_shape,
                                     self._mesh_sizes(path[0].device_mesh)) if path else 0.0
        sizes = self._mesh_sizes(path[0].device_mesh)
        elems = [self._local_elems(_state_of(l.sharding_spec), l.global_shape, sizes) for l in path]
        return max(a + b for a, b in zip(elems[:-1], elems[1:]))

    def apply(self, tensor: torch.Tensor, source: Layout, target: Layout) -> torch.Tensor:
        _, comms = self.layout_converting(source, target)
        for cs in comms:
            tensor = cs.covert_spec_to_action(tensor)
        return tensor
