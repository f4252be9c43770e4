"""`symbolic_trace(model, meta_args=...)`: torch.fx tracing with concrete (meta) example inputs so data-dependent
shapes resolve.  Parity: reference `colossalai/fx/tracer/tracer.py` (`ColoTracer`) / `_analyzer/fx/tracer`."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.fx as fx
import torch.nn as nn

__all__ = ["ColoTracer", "ColoGraphModule", "symbolic_trace"]


class ColoGraphModule(fx.GraphModule):
    """GraphModule that remembers the meta arguments it was traced with."""

    def __init__(self, root, graph, class_name: str = "GraphModule", meta_args: Optional[Dict[str, Any]] = None):
        super().__init__(root, graph, class_name)
        self.meta_args = meta_args or {}


class ColoTracer(fx.Tracer):
    """Leaf policy: torch.nn leaves stay leaves; user modules are traced through."""

    def __init__(self, leaf_modules=(), **kw) -> None:
        super().__init__(**kw)
        self._extra_leaves = tuple(leaf_modules)

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        if isinstance(m, self._extra_leaves):
            return True
        return super().is_leaf_module(m, module_qualified_name)


def symbolic_trace(root: nn.Module, concrete_args: Optional[Dict[str, Any]] = None,
                   meta_args: Optional[Dict[str, Any]] = None, leaf_modules=()) -> ColoGraphModule:
    """Trace `root`; arguments listed in `concrete_args` are baked in, every other argument stays a placeholder.
    `meta_args` (name -> meta tensor) are recorded for `MetaInfoProp`."""
    tracer = ColoTracer(leaf_modules=leaf_modules)
    graph = tracer.trace(root, concrete_args=concrete_args)
    return ColoGraphModule(tracer.root, graph, root.__class__.__name__, meta_args=meta_args)
