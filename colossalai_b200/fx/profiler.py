"""Shape / FLOP / memory annotation of a traced graph.  Parity: reference `colossalai/fx/passes/meta_info_prop.py`
(`MetaInfoProp`) and `fx/profiler` (per-node fwd flop + activation size)."""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch
import torch.fx as fx
from torch.utils._pytree import tree_map
from torch.utils.flop_counter import FlopCounterMode

__all__ = ["MetaInfoProp", "profile_flops_and_memory"]


def _nbytes(x) -> int:
    total = 0

    def f(t):
        nonlocal total
        if torch.is_tensor(t):
            total += t.numel() * t.element_size()
        return t

    tree_map(f, x)
    return total


class MetaInfoProp(fx.Interpreter):
    """Runs the graph on meta tensors and records per node: `meta['tensor_meta']` (shape/dtype), `meta['fwd_flop']`,
    `meta['fwd_out_bytes']`."""

    def run_node(self, n: fx.Node) -> Any:
        with FlopCounterMode(display=False) as fc:
            out = super().run_node(n)
        n.meta["fwd_flop"] = fc.get_total_flops()
        n.meta["fwd_out_bytes"] = _nbytes(out)
        n.meta["tensor_meta"] = tree_map(lambda t: (tuple(t.shape), t.dtype) if torch.is_tensor(t) else t, out)
        return out

    def propagate(self, *args):
        params = list(self.module.parameters())
        on_meta = bool(params) and all(p.device.type == "meta" for p in params)
        # a meta model propagates on meta inputs (no memory, no compute); a materialised one on the given inputs
        margs = tree_map(lambda t: t.to("meta") if (torch.is_tensor(t) and on_meta) else t, args)
        return super().run(*margs)

    def summary(self) -> str:
        rows = ["node                           op             GFLOP     out MB"]
        for n in self.module.graph.nodes:
            rows.append(f"{n.name[:30]:30s} {n.op[:14]:14s} {n.meta.get('fwd_flop', 0) / 1e9:8.3f} "
                        f"{n.meta.get('fwd_out_bytes', 0) / 2**20:10.2f}")
        return "\n".join(rows)


def profile_flops_and_memory(module: torch.nn.Module, *example_inputs) -> Tuple[int, int]:
    """(total forward FLOPs, total activation bytes) of `module` on meta copies of `example_inputs`."""
    import copy

    m = copy.deepcopy(module).to("meta")
    args = tree_map(lambda t: t.to("meta") if torch.is_tensor(t) else t, example_inputs)
    acts = 0

    def hook(_m, _i, out):
        nonlocal acts
        acts += _nbytes(out)

    hs = [sub.register_forward_hook(hook) for sub in m.modules() if not list(sub.children())]
    with FlopCounterMode(display=False) as fc:
        m(*args)
    for h in hs:
        h.remove()
    return fc.get_total_flops(), acts
