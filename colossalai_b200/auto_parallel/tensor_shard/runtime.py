"""Apply a solved plan to the traced module: shard parameters, insert the layout conversions and the partial-sum
reductions as graph nodes, split the inputs, rewrite reshape constants.

Parity: reference `colossalai/auto_parallel/passes/{runtime_preparation_pass,runtime_apply_pass}.py` (parameter
sharding + `runtime_apply` / `runtime_comm_spec_apply` nodes).  Every inserted node is a plain module-level function of
this file so the generated `forward` stays importable; the mesh is referenced through a small registry id because fx
arguments must be literals.
"""
from __future__ import annotations

import operator
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

from ...device import DeviceMesh
from .node_handler import node_shape, tensor_operands
from .sharding_strategy import ShardingStrategy, Spec, replicated
from .solver import resharding_steps

__all__ = ["reshard", "reduce_fwd", "reduce_bwd", "global_size", "ParallelLinear", "ParallelEmbedding", "shard_tensor",
           "runtime_apply_pass", "register_mesh"]

_MESHES: Dict[int, DeviceMesh] = {}


def register_mesh(mesh: DeviceMesh) -> int:
    for k, m in _MESHES.items():
        if m is mesh:
            return k
    _MESHES[len(_MESHES)] = mesh
    return len(_MESHES) - 1


def _group(mesh_id: int, axis: int):
    return _MESHES[mesh_id].get_process_group(axis)


# ---- graph-level communication nodes (autograd aware)
def reshard(x: torch.Tensor, src: Spec, dst: Spec, mesh_id: int) -> torch.Tensor:
    from ...shardformer.layer._operation import (all_to_all_comm, gather_forward_split_backward,
                                                 split_forward_gather_backward)

    mesh = _MESHES[mesh_id]
    for kind, a, sd, dd in resharding_steps(tuple(src), tuple(dst), tuple(mesh.shape)):
        g = mesh.get_process_group(a)
        if kind == "gather":
            x = gather_forward_split_backward(x, sd, g)
        elif kind == "split":
            x = split_forward_gather_backward(x, dd, g)
        else:
            x = all_to_all_comm(x, g, scatter_dim=dd, gather_dim=sd)
    return x


def reduce_fwd(x: torch.Tensor, axes: Sequence[int], mesh_id: int) -> torch.Tensor:
    """Partial sums -> full value (forward all-reduce, identity backward)."""
    from ...shardformer.layer._operation import reduce_forward

    for a in axes:
        x = reduce_forward(x, _group(mesh_id, a))
    return x


def reduce_bwd(x: torch.Tensor, axes: Sequence[int], mesh_id: int) -> torch.Tensor:
    """Replicated operand of a sharded computation: identity forward, all-reduce of its gradient backward."""
    from ...shardformer.layer._operation import reduce_backward

    if not (torch.is_tensor(x) and (x.requires_grad or torch.is_grad_enabled())):
        return x
    for a in axes:
        x = reduce_backward(x, _group(mesh_id, a))
    return x


def global_size(size, spec: Sequence[Optional[int]], mesh_shape: Sequence[int], dim: Optional[int] = None):
    """Local `x.size()` / `x.shape` / `x.size(dim)` of a sharded tensor -> the extents of the whole tensor."""
    if dim is not None:
        a = spec[dim % len(spec)]
        return size if a is None else size * mesh_shape[a]
    return torch.Size([s if a is None else s * mesh_shape[a] for s, a in zip(size, spec)])


# ---- parameters
def shard_tensor(t: torch.Tensor, spec: Sequence[Optional[int]], mesh: DeviceMesh) -> torch.Tensor:
    rank = dist.get_rank() if dist.is_initialized() else 0
    coord = mesh.global_rank_to_local_rank(rank)
    for d, a in enumerate(spec):
        if a is None:
            continue
        t = t.chunk(mesh.shape[a], dim=d)[coord[a]]
    return t.contiguous()


def _sync_grads(params, axes: Sequence[int], mesh_id: int) -> None:
    """Data-parallel style synchronisation: the parameter is replicated along `axes` while its node saw different
    samples there."""
    if not axes:
        return
    for p in params:
        if p is None or not p.requires_grad:
            continue
        done = getattr(p, "_autoshard_sync_axes", set())
        for a in axes:
            if a in done:
                continue
            done.add(a)

            def hook(g, a=a):
                g = g.contiguous().clone()
                dist.all_reduce(g, group=_group(mesh_id, a))
                return g

            p.register_hook(hook)
        p._autoshard_sync_axes = done


class ParallelLinear(nn.Module):
    """`nn.Linear` with the weight laid out by a solved strategy: `[out, in]` split over (col axis, row axis); the
    row-parallel partial sum is all-reduced before the bias is added."""

    def __init__(self, lin: nn.Linear, strategy: ShardingStrategy, mesh: DeviceMesh, mesh_id: int) -> None:
        super().__init__()
        wspec = strategy.param_specs["weight"]
        self.in_features, self.out_features = lin.in_features, lin.out_features
        self.weight = nn.Parameter(shard_tensor(lin.weight.data, wspec, mesh), requires_grad=lin.weight.requires_grad)
        self.bias = None
        if lin.bias is not None:
            self.bias = nn.Parameter(shard_tensor(lin.bias.data, strategy.param_specs["bias"], mesh),
                                     requires_grad=lin.bias.requires_grad)
        self.reduce_axes = tuple(strategy.reduce_axes)
        self.mesh_id = mesh_id
        self.weight_spec = tuple(wspec)
        _sync_grads([self.weight, self.bias], strategy.grad_sync_axes, mesh_id)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = F.linear(x, self.weight)
        if self.reduce_axes:
            y = reduce_fwd(y, self.reduce_axes, self.mesh_id)
        return y if self.bias is None else y + self.bias

    def extra_repr(self) -> str:
        return f"in={self.in_features}, out={self.out_features}, weight_spec={self.weight_spec}"


class ParallelEmbedding(nn.Module):
    """`nn.Embedding` with the hidden dimension split (no communication: every rank looks up its own columns)."""

    def __init__(self, emb: nn.Embedding, strategy: ShardingStrategy, mesh: DeviceMesh, mesh_id: int) -> None:
        super().__init__()
        self.num_embeddings, self.embedding_dim = emb.num_embeddings, emb.embedding_dim
        self.padding_idx = emb.padding_idx
        self.weight = nn.Parameter(shard_tensor(emb.weight.data, strategy.param_specs["weight"], mesh),
                                   requires_grad=emb.weight.requires_grad)
        _sync_grads([self.weight], strategy.grad_sync_axes, mesh_id)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return F.embedding(ids, self.weight, self.padding_idx)


def _set_submodule(gm: fx.GraphModule, target: str, new: nn.Module) -> None:
    parent_name, _, child = target.rpartition(".")
    parent = gm.get_submodule(parent_name) if parent_name else gm
    setattr(parent, child, new)


def _rewrite_reshape_constants(node: fx.Node, out_spec: Spec, mesh_shape) -> None:
    if not any(a is not None for a in out_spec):
        return
    name = node.target if isinstance(node.target, str) else getattr(node.target, "__name__", "")
    if name not in ("view", "reshape"):
        return
    args = list(node.args)
    packed = len(args) == 2 and isinstance(args[1], (tuple, list, torch.Size))
    dims = list(args[1]) if packed else args[1:]
    if len(dims) != len(out_spec):
        return
    for j, a in enumerate(out_spec):
        if a is None:
            continue
        if isinstance(dims[j], int) and dims[j] > 0:
            dims[j] = dims[j] // mesh_shape[a]
        elif isinstance(dims[j], fx.Node):                # a (global) extent computed in the graph
            with node.graph.inserting_before(node):
                dims[j] = node.graph.call_function(operator.floordiv, (dims[j], mesh_shape[a]))
    node.args = (args[0], tuple(dims)) if packed else (args[0], *dims)


def runtime_apply_pass(gm: fx.GraphModule, solution: Dict[fx.Node, ShardingStrategy], mesh: DeviceMesh
                       ) -> Tuple[fx.GraphModule, Dict[str, Spec]]:
    """Returns the transformed module and `{parameter name: spec}` for the parameters that were sharded."""
    assert mesh.is_initialized, "the device mesh needs its process groups (DeviceMesh(..., init_process_group=True))"
    mesh_id = register_mesh(mesh)
    ms = tuple(mesh.shape)
    graph = gm.graph
    param_specs: Dict[str, Spec] = {}
    produced: Dict[fx.Node, fx.Node] = {}            # original node -> node that carries its solved layout
    cache: Dict[Tuple[fx.Node, Spec, Tuple[int, ...]], fx.Node] = {}

    for node in list(graph.nodes):
        s = solution.get(node)
        if s is None:
            continue
        # ---- operands: layout conversion + backward reductions
        for u in tensor_operands(node):
            want = s.input_specs.get(u.name)
            su = solution.get(u)
            if want is None or su is None or su.output_spec is None:
                continue
            src_node = produced.get(u, u)
            bwd_axes = tuple(s.bwd_reduce_inputs.get(u.name, ()))
            key = (u, tuple(want), bwd_axes)
            if key not in cache:
                cur = src_node
                with graph.inserting_before(node):
                    if tuple(su.output_spec) != tuple(want):
                        cur = graph.call_function(reshard, (cur, tuple(su.output_spec), tuple(want), mesh_id))
                    if bwd_axes:
                        cur = graph.call_function(reduce_bwd, (cur, bwd_axes, mesh_id))
                cache[key] = cur
            new_in = cache[key]
            if new_in is not u:
                node.replace_input_with(u, new_in)
        # ---- the node itself
        if node.op == "placeholder":
            if s.output_spec is not None and any(a is not None for a in s.output_spec):
                with graph.inserting_after(node):
                    split = graph.call_function(reshard, (node, replicated(len(s.output_spec)), tuple(s.output_spec),
                                                          mesh_id))
                produced[node] = split
        elif node.op == "call_module":
            mod = gm.get_submodule(node.target)
            if type(mod) is nn.Linear:
                _set_submodule(gm, node.target, ParallelLinear(mod, s, mesh, mesh_id))
            elif type(mod) is nn.Embedding and "weight" in s.param_specs:
                _set_submodule(gm, node.target, ParallelEmbedding(mod, s, mesh, mesh_id))
            else:
                _sync_grads(list(mod.parameters(recurse=False)), s.grad_sync_axes, mesh_id)
            for pname, spec in s.param_specs.items():
                if any(a is not None for a in spec):
                    param_specs[f"{node.target}.{pname}"] = tuple(spec)
        else:
            is_size = (node.op == "call_method" and node.target == "size") or \
                (node.op == "call_function" and node.target is getattr and len(node.args) == 2 and node.args[1] == "shape")
            if is_size and len(s.input_specs) == 1:
                xspec = tuple(next(iter(s.input_specs.values())))
                if any(a is not None for a in xspec):
                    dim = node.args[1] if (node.target == "size" and len(node.args) > 1) else node.kwargs.get("dim")
                    with graph.inserting_after(node):
                        produced[node] = graph.call_function(global_size, (node, xspec, ms, dim))
            if s.output_spec is not None:
                _rewrite_reshape_constants(node, s.output_spec, ms)
            if s.reduce_axes:
                with graph.inserting_after(node):
                    red = graph.call_function(reduce_fwd, (node, tuple(s.reduce_axes), mesh_id))
                produced[node] = red
    # users of a node that now has a successor carrying its value (split placeholder / reduced matmul)
    for orig, new in produced.items():
        for user in list(orig.users):
            if user is new:
                continue
            user.replace_input_with(orig, new)
    graph.lint()
    gm.recompile()
    return gm, param_specs
