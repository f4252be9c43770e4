"""CUDA-graph capture of the decode step per batch size.
Parity: reference `colossalai/inference/graph_runner.py:9-100` (capture once with static buffers, replay with copy-in)."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

__all__ = ["CUDAGraphRunner"]


class CUDAGraphRunner:
    def __init__(self, fn: Callable) -> None:
        self.fn = fn
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_inputs: Dict[str, torch.Tensor] = {}
        self.static_output: Optional[torch.Tensor] = None

    def capture(self, memory_pool=None, stream=None, **tensors: torch.Tensor) -> None:
        assert self.graph is None, "graph already captured"
        self.static_inputs = {k: v.clone() for k, v in tensors.items()}
        # one eager warm-up so lazy allocations / autotuning happen outside the capture
        self.fn(**self.static_inputs)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=memory_pool, stream=stream):
            self.static_output = self.fn(**self.static_inputs)
        torch.cuda.synchronize()

    def forward(self, **tensors: torch.Tensor) -> torch.Tensor:
        for k, v in tensors.items():
            self.static_inputs[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_output

    __call__ = forward
