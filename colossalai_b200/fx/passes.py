"""Graph passes on `torch.fx` graphs: pipeline splitting (balanced by parameter count or uniform by node count) and
activation-checkpoint region rewriting.

Parity: reference `colossalai/fx/passes/{adding_split_node_pass.py (balanced_split_pass, uniform_split_pass,
split_with_split_nodes_pass), split_module.py, meta_info_prop.py}` and `fx/codegen/activation_checkpoint_codegen.py`
(annotated `activation_checkpoint` regions emitted as `torch.utils.checkpoint` calls)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.fx as fx
import torch.nn as nn
from torch.fx.passes.split_module import split_module
from torch.utils.checkpoint import checkpoint as torch_checkpoint

__all__ = ["balanced_split_pass", "uniform_split_pass", "split_with_split_nodes_pass",
           "activation_checkpoint_pass", "CheckpointRegion"]


def _param_count(gm: fx.GraphModule, node: fx.Node) -> int:
    if node.op == "call_module":
        return sum(p.numel() for p in gm.get_submodule(node.target).parameters())
    if node.op == "get_attr":
        t = gm
        for a in node.target.split("."):
            t = getattr(t, a)
        return t.numel() if isinstance(t, torch.Tensor) else 0
    return 0


def _assign(gm: fx.GraphModule, boundaries: List[fx.Node]) -> Dict[fx.Node, int]:
    part, cur = {}, 0
    cut = set(boundaries)
    for n in gm.graph.nodes:
        part[n] = cur
        if n in cut:
            cur += 1
    return part


def balanced_split_pass(gm: fx.GraphModule, pp_size: int) -> Dict[fx.Node, int]:
    """Cut after the node where the running parameter count crosses each 1/pp_size share."""
    nodes = [n for n in gm.graph.nodes if n.op not in ("placeholder", "output")]
    total = sum(_param_count(gm, n) for n in nodes)
    share, acc, cuts = total / pp_size, 0, []
    for n in nodes:
        acc += _param_count(gm, n)
        if len(cuts) < pp_size - 1 and acc >= share * (len(cuts) + 1):
            cuts.append(n)
    return _assign(gm, cuts)


def uniform_split_pass(gm: fx.GraphModule, pp_size: int) -> Dict[fx.Node, int]:
    """Equal number of parameter-carrying nodes per stage."""
    nodes = [n for n in gm.graph.nodes if n.op == "call_module" and _param_count(gm, n) > 0]
    per = max(1, len(nodes) // pp_size)
    cuts = [nodes[(i + 1) * per - 1] for i in range(pp_size - 1) if (i + 1) * per - 1 < len(nodes)]
    return _assign(gm, cuts)


def split_with_split_nodes_pass(gm: fx.GraphModule, partition: Dict[fx.Node, int]) -> Tuple[fx.GraphModule, List[nn.Module]]:
    """Materialise the stage sub-modules (`submod_0..k`) of a partition."""
    split = split_module(gm, gm, lambda n: partition[n])
    stages = [m for name, m in split.named_children() if name.startswith("submod_")]
    return split, stages


class CheckpointRegion(nn.Module):
    def __init__(self, gm: fx.GraphModule) -> None:
        super().__init__()
        self.gm = gm

    def forward(self, *args):
        if self.training and torch.is_grad_enabled():
            return torch_checkpoint(self.gm, *args, use_reentrant=False)
        return self.gm(*args)


def activation_checkpoint_pass(gm: fx.GraphModule, regions: Sequence[Sequence[str]]) -> fx.GraphModule:
    """`regions`: lists of node NAMES; every list becomes one recomputed segment.  Implemented by splitting the graph
    at the region boundaries and wrapping the region sub-graphs in `CheckpointRegion`."""
    region_of: Dict[str, int] = {name: i for i, names in enumerate(regions) for name in names}
    ids: Dict[fx.Node, int] = {}
    cur, last_key = 0, None
    for n in gm.graph.nodes:
        key = ("r", region_of[n.name]) if n.name in region_of else ("n",)
        if n.op in ("placeholder", "output"):
            ids[n] = cur
            continue
        if last_key is not None and key != last_key:
            cur += 1
        ids[n] = cur
        last_key = key
    is_region = {}
    for n, i in ids.items():
        if n.name in region_of:
            is_region[i] = True
    split = split_module(gm, gm, lambda n: ids[n])
    for name, child in list(split.named_children()):
        if name.startswith("submod_") and is_region.get(int(name.split("_")[1]), False):
            setattr(split, name, CheckpointRegion(child))
    return split
