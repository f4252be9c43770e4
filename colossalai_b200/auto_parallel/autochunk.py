"""AutoChunk: cut peak activation memory by evaluating token-wise regions of a model in slices along one dimension.

The reference traces the model with fx, estimates per-node activation memory, searches "chunk regions" whose inputs
and outputs can be sliced along a common dim without changing the result, and regenerates the forward code with loops
(`colossalai/autochunk/{autochunk_codegen.py, search_chunk.py, trace_flow.py, trace_indice.py, estimate_memory.py}`,
~4 k lines).  Here the same idea is applied at module granularity: a region is a sub-module that is token-wise along
`dim` (verified numerically on a probe), its peak memory is measured with meta tensors (`fx.MetaInfoProp`-style
accounting through saved-tensor hooks), and `autochunk(model, max_memory_bytes)` wraps exactly the regions that
exceed the budget with the smallest chunk count that fits.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn as nn

__all__ = ["ChunkedModule", "is_chunkable", "estimate_activation_bytes", "search_chunk_regions", "autochunk"]


class ChunkedModule(nn.Module):
    """Run `module` on `n_chunks` slices of the input along `dim` and concatenate (exact for token-wise modules)."""

    def __init__(self, module: nn.Module, dim: int, n_chunks: int) -> None:
        super().__init__()
        self.module, self.dim, self.n_chunks = module, dim, n_chunks

    def forward(self, x: torch.Tensor, *args, **kwargs):
        if self.n_chunks <= 1 or x.shape[self.dim] < self.n_chunks:
            return self.module(x, *args, **kwargs)
        outs = [self.module(c, *args, **kwargs) for c in x.chunk(self.n_chunks, dim=self.dim)]
        if isinstance(outs[0], tuple):
            return tuple(torch.cat([o[i] for o in outs], dim=self.dim) for i in range(len(outs[0])))
        return torch.cat(outs, dim=self.dim)


@torch.no_grad()
def is_chunkable(module: nn.Module, example: torch.Tensor, dim: int, atol: float = 1e-5) -> bool:
    """Numerical check that slicing along `dim` commutes with the module (no cross-token mixing)."""
    if example.shape[dim] < 2:
        return False
    was = module.training
    module.eval()
    try:
        full = module(example)
        half = torch.cat([module(c) for c in example.chunk(2, dim=dim)], dim=dim)
        ok = isinstance(full, torch.Tensor) and full.shape == half.shape and torch.allclose(full, half, atol=atol,
                                                                                           rtol=1e-4)
    except Exception:
        ok = False
    module.train(was)
    return bool(ok)


def estimate_activation_bytes(module: nn.Module, example: torch.Tensor) -> int:
    """Bytes autograd keeps alive for the module's backward + its output (a proxy of the node-level peak)."""
    saved = [0]

    def pack(t):
        saved[0] += t.numel() * t.element_size()
        return t

    x = example.detach().requires_grad_(example.is_floating_point())
    with torch.autograd.graph.saved_tensors_hooks(pack, lambda t: t):
        y = module(x)
    y = y[0] if isinstance(y, tuple) else y
    return int(saved[0] + y.numel() * y.element_size())


def search_chunk_regions(model: nn.Module, probes: Dict[str, torch.Tensor], dim: int,
                         max_memory_bytes: float) -> List[Tuple[str, int, int]]:
    """`probes`: {module path: example input}.  Returns [(path, bytes, n_chunks)] for chunkable regions over budget."""
    named = dict(model.named_modules())
    out = []
    for path, ex in probes.items():
        m = named[path]
        b = estimate_activation_bytes(m, ex)
        if b > max_memory_bytes and is_chunkable(m, ex, dim):
            n = min(ex.shape[dim], 2 ** math.ceil(math.log2(b / max_memory_bytes)))
            out.append((path, b, int(n)))
    return out


def autochunk(model: nn.Module, example_kwargs: Dict[str, torch.Tensor], max_memory_bytes: float, dim: int = 0,
              target_types: Optional[Iterable[type]] = None) -> Tuple[nn.Module, List[Tuple[str, int, int]]]:
    """Capture the inputs of every candidate sub-module on one probe forward, pick the regions whose activation
    footprint exceeds `max_memory_bytes`, and wrap them in `ChunkedModule`s."""
    if target_types is None:
        from ..models.transformer import MLP

        target_types = (MLP,)
    target_types = tuple(target_types)
    probes: Dict[str, torch.Tensor] = {}
    hooks = []
    for name, m in model.named_modules():
        if isinstance(m, target_types):
            hooks.append(m.register_forward_pre_hook(
                lambda mod, args, name=name: probes.__setitem__(name, args[0].detach())))
    was = model.training
    model.eval()
    with torch.no_grad():
        model(**example_kwargs)
    model.train(was)
    for h in hooks:
        h.remove()
    regions = search_chunk_regions(model, probes, dim, max_memory_bytes)
    for path, _, n in regions:
        parent = model
        *pp, leaf = path.split(".")
        for p in pp:
            parent = getattr(parent, p)
        setattr(parent, leaf, ChunkedModule(getattr(parent, leaf), dim, n))
    return model, regions
