"""`symbolic_profile(gm, *args)`: run a traced module node by node on meta tensors and attach a `MetaInfo` to every
node — forward / backward FLOPs, bytes produced, bytes autograd saves for backward, parameter bytes.

Parity: reference `colossalai/_analyzer/fx/symbolic_profile.py` (`symbolic_profile`, `register_flop_count_impl`) +
`fx/passes/shape_prop.py`.  Backward cost of a node is measured in isolation: its inputs are detached leaves, the node
is re-run under the FLOP interception mode and differentiated on the spot; `saved_tensors_hooks` reports what autograd
holds on to."""
from __future__ import annotations

from typing import Any, Dict

import torch
import torch.fx as fx
from torch.utils._pytree import tree_map

from .._subclasses.flop_tensor import Phase, _FlopMode
from .node_util import MetaInfo

__all__ = ["symbolic_profile"]


def _nbytes(x) -> int:
    total = 0

    def f(t):
        nonlocal total
        if isinstance(t, torch.Tensor):
            total += t.numel() * t.element_size()
        return t

    tree_map(f, x)
    return total


class _Profiler(fx.Interpreter):
    def run_node(self, n: fx.Node) -> Any:
        info = MetaInfo()
        if n.op in ("placeholder", "get_attr", "output"):
            out = super().run_node(n)
            info.outputs = tree_map(lambda t: (tuple(t.shape), t.dtype) if isinstance(t, torch.Tensor) else t, out)
            info.output_bytes = _nbytes(out) if n.op != "output" else 0
            n.meta["info"] = info
            return out
        args, kwargs = self.fetch_args_kwargs_from_env(n)
        # isolate the node: floating inputs become fresh leaves so backward stops at the node boundary
        leaves = []

        def leaf(t):
            if isinstance(t, torch.Tensor) and t.is_floating_point():
                t = t.detach().requires_grad_(True)
                leaves.append(t)
            return t

        largs, lkwargs = tree_map(leaf, args), tree_map(leaf, kwargs)
        saved: Dict[int, int] = {}

        def pack(t):
            saved[id(t)] = t.numel() * t.element_size()
            return t

        mode = _FlopMode()
        with mode, torch.autograd.graph.saved_tensors_hooks(pack, lambda t: t):
            out = getattr(self, n.op)(n.target, largs, lkwargs)
            mode.phase = Phase.BWD
            outs = []
            tree_map(lambda t: outs.append(t) if isinstance(t, torch.Tensor) and t.requires_grad else None, out)
            if outs:
                torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
        info.fwd_flop, info.bwd_flop = mode.flops[Phase.FWD], mode.flops[Phase.BWD]
        info.output_bytes = _nbytes(out)
        info.saved_bytes = sum(saved.values())
        if n.op == "call_module":
            info.param_bytes = sum(p.numel() * p.element_size() for p in self.fetch_attr(n.target).parameters())
        info.outputs = tree_map(lambda t: (tuple(t.shape), t.dtype) if isinstance(t, torch.Tensor) else t, out)
        n.meta["info"] = info
        n.meta.setdefault("tensor_meta", info.outputs)
        return tree_map(lambda t: t.detach() if isinstance(t, torch.Tensor) else t, out)


def symbolic_profile(gm: fx.GraphModule, *args, verbose: bool = False) -> fx.GraphModule:
    """Annotates `gm` in place (`node.meta['info']`) and returns it.  Arguments may live anywhere: they are moved to
    the meta device together with a meta copy of the parameters, so nothing is computed or allocated."""
    import copy

    meta_gm = copy.deepcopy(gm).to("meta")
    margs = tree_map(lambda t: t.detach().to("meta") if isinstance(t, torch.Tensor) else t, args)
    _Profiler(meta_gm).run(*margs)
    for src, dst in zip(meta_gm.graph.nodes, gm.graph.nodes):
        dst.meta["info"] = src.meta.get("info", MetaInfo())
        if "tensor_meta" in src.meta:
            dst.meta.setdefault("tensor_meta", src.meta["tensor_meta"])
    if verbose:
        print(f"{'node':28s} {'fwd GFLOP':>10s} {'bwd GFLOP':>10s} {'out MB':>9s} {'saved MB':>9s}")
        for n in gm.graph.nodes:
            i = n.meta["info"]
            print(f"{n.name[:28]:28s} {i.fwd_flop / 1e9:10.3f} {i.bwd_flop / 1e9:10.3f} {i.output_bytes / 2**20:9.2f} "
                  f"{i.saved_bytes / 2**20:9.2f}")
    return gm
