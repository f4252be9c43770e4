"""`with PipelinableContext(): model = Net()` records every layer in construction order; `to_layer_list()` fixes the
execution order and `partition(num_chunks, pipeline_size, rank)` builds this rank's stage module.
Parity: reference `colossalai/legacy/pipeline/pipelinable.py:1-260`, `layer_spec.py`, `utils.py:partition_*`."""
from __future__ import annotations

import heapq
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ...utils.model.utils import InsertPostInitMethodToModuleSubClasses

__all__ = ["LayerSpec", "PipelinableContext", "PipelinableModel", "partition_uniform", "partition_balanced"]


class LayerSpec:
    """Deferred construction of one layer: `LayerSpec(nn.Linear, 4, 8).build()`."""

    def __init__(self, typename: type, *module_args, **module_kwargs) -> None:
        self.typename, self.module_args, self.module_kwargs = typename, module_args, module_kwargs
        self.children = None
        self._param_count = 0

    def build(self) -> nn.Module:
        return self.typename(*self.module_args, **self.module_kwargs)

    def set_children(self, children) -> None:
        self.children = children

    def count_params(self) -> int:
        if self._param_count == 0:
            with torch.device("meta"):
                m = self.build()
            self._param_count = sum(p.numel() for p in m.parameters())
        return self._param_count


def partition_uniform(num_items: int, pipeline_parallel_size: int, num_chunks: int = 1) -> List[List[Tuple[int, int]]]:
    """parts[rank] = [(start, end), ...] one range per model chunk (interleaved assignment)."""
    assert num_items % num_chunks == 0 or num_chunks == 1
    parts: List[List[Tuple[int, int]]] = [[] for _ in range(pipeline_parallel_size)]
    per_chunk = num_items // num_chunks
    for c in range(num_chunks):
        base = c * per_chunk
        size, left = divmod(per_chunk if c < num_chunks - 1 else num_items - base, pipeline_parallel_size)
        start = base
        for r in range(pipeline_parallel_size):
            n = size + (1 if r < left else 0)
            parts[r].append((start, start + n))
            start += n
    return parts


def partition_balanced(weights: Sequence[float], pipeline_parallel_size: int, num_chunks: int = 1
                       ) -> List[List[Tuple[int, int]]]:
    """Contiguous ranges whose weight sums are as even as possible (binary search on the bottleneck)."""
    n, k = len(weights), pipeline_parallel_size * num_chunks
    if n <= k:
        return partition_uniform(n, pipeline_parallel_size, num_chunks)

    def cuts_for(limit: float) -> Optional[List[int]]:
        cuts, acc = [], 0.0
        for i, w in enumerate(weights):
            if w > limit:
                return None
            if acc + w > limit:
                cuts.append(i)
                acc = 0.0
            acc += w
        return cuts if len(cuts) <= k - 1 else None

    lo, hi = max(weights), sum(weights)
    for _ in range(50):
        mid = (lo + hi) / 2
        if cuts_for(mid) is None:
            lo = mid
        else:
            hi = mid
    cuts = cuts_for(hi) or []
    bounds = [0] + cuts + [n]
    while len(bounds) < k + 1:            # fewer ranges than stages: split the largest ranges
        i = max(range(len(bounds) - 1), key=lambda j: bounds[j + 1] - bounds[j])
        if bounds[i + 1] - bounds[i] < 2:
            break
        bounds.insert(i + 1, (bounds[i] + bounds[i + 1]) // 2)
    ranges = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
    ranges += [(n, n)] * (k - len(ranges))
    parts: List[List[Tuple[int, int]]] = [[] for _ in range(pipeline_parallel_size)]
    for i, r in enumerate(ranges):
        parts[i % pipeline_parallel_size].append(r)
    return parts


class PipelinableModel(nn.Module):
    """This rank's slice of the layer list; non-module callables recorded in the exec sequence are kept in order."""

    def __init__(self, module_list: List[Any], front_func_dict: Dict, behind_func_dict: Dict) -> None:
        super().__init__()
        self._module_list = nn.ModuleList([m for m in module_list if isinstance(m, nn.Module)])
        self._order = module_list
        self._front, self._behind = front_func_dict, behind_func_dict

    def forward(self, *args, **kwargs):
        x = args[0] if len(args) == 1 else args
        for m in self._order:
            mid = id(m)
            for f in self._front.get(mid, []):
                x = f(x)
            x = m(*x) if isinstance(x, tuple) else m(x)
            for f in self._behind.get(mid, []):
                x = f(x)
        return x


class PipelinableContext(InsertPostInitMethodToModuleSubClasses):
    def __init__(self, policy: str = "balanced") -> None:
        super().__init__()
        self._layer_spec_dict: Dict[int, LayerSpec] = {}
        self._root_children = None
        self._model: Optional[nn.Module] = None
        self._layer_spec_list: List[LayerSpec] = []
        self._func_dict: Dict[int, Callable] = {}
        self._policy = policy
        self._modules_in_order: List[nn.Module] = []
        self._front: Dict[int, List[Callable]] = {}
        self._behind: Dict[int, List[Callable]] = {}

    @property
    def policy(self) -> str:
        return self._policy

    @policy.setter
    def policy(self, p: str) -> None:
        self._policy = p

    @property
    def layers_count(self) -> int:
        return len(self._modules_in_order)

    def _pre_context_exec(self) -> None:
        pass

    def _post_context_exec(self) -> None:
        pass

    def _post_init_method(self, module: nn.Module, *args, **kwargs) -> None:
        self._layer_spec_dict[id(module)] = LayerSpec(type(module), *args, **kwargs)
        self._model = module            # the last module constructed is the root

    def to_layer_list(self, exec_seq: Optional[List] = None) -> None:
        """`exec_seq`: names of the root's children (strings), callables to run between them, or None for
        registration order of the root's direct children."""
        assert self._model is not None, "construct the model inside the context first"
        children = dict(self._model.named_children())
        self._modules_in_order, self._front, self._behind = [], {}, {}
        if exec_seq is None:
            self._modules_in_order = list(children.values())
            return
        pending_front: List[Callable] = []
        for item in exec_seq:
            if isinstance(item, str):
                if item == "front" or item == "behind":
                    continue
                m = children[item]
                self._modules_in_order.append(m)
                if pending_front:
                    self._front[id(m)] = pending_front
                    pending_front = []
            elif isinstance(item, tuple) and callable(item[0]):        # (func, "front"|"behind")
                if item[1] == "front":
                    pending_front.append(item[0])
                else:
                    self._behind.setdefault(id(self._modules_in_order[-1]), []).append(item[0])
            elif callable(item):
                if self._modules_in_order:
                    self._behind.setdefault(id(self._modules_in_order[-1]), []).append(item)
                else:
                    pending_front.append(item)

    def partition(self, num_chunks: int, pipeline_size: int, rank: int):
        mods = self._modules_in_order
        if self._policy == "uniform":
            parts = partition_uniform(len(mods), pipeline_size, num_chunks)[rank]
        else:
            w = [max(1, sum(p.numel() for p in m.parameters())) for m in mods]
            parts = partition_balanced(w, pipeline_size, num_chunks)[rank]
        models = [PipelinableModel(mods[s:e], self._front, self._behind) for s, e in parts]
        return models[0] if num_chunks == 1 else nn.ModuleList(models)
