"""`ProcessGroup(rank, ranks, tp_degree, dp_degree)`: the world as a [dp, tp] grid, tensor-parallel ranks adjacent
(reference `legacy/tensor/process_group.py:37-330`).  Torch groups are created once per rank list and shared."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch.distributed as dist

__all__ = ["ProcessGroup"]

_PG_CACHE: Dict[Tuple[Tuple[int, ...], str], object] = {}


def _get_group(ranks: List[int], backend: Optional[str]):
    backend = backend or dist.get_backend()
    key = (tuple(ranks), backend)
    if key not in _PG_CACHE:
        _PG_CACHE[key] = dist.new_group(list(ranks), backend=backend)
    return _PG_CACHE[key]


class ProcessGroup:
    def __init__(self, rank: Optional[int] = None, ranks: Optional[List[int]] = None, tp_degree: Optional[int] = None,
                 dp_degree: Optional[int] = None, backend: Optional[str] = None) -> None:
        if not dist.is_initialized():
            self.is_init = False
            return
        self._rank = dist.get_rank() if rank is None else rank
        self._rank_list = list(range(dist.get_world_size())) if ranks is None else sorted(ranks)
        world = len(self._rank_list)
        if tp_degree is None and dp_degree is None:
            dp_degree, tp_degree = world, 1
        elif tp_degree is None:
            tp_degree = world // dp_degree
        elif dp_degree is None:
            dp_degree = world // tp_degree
        assert tp_degree * dp_degree == world, f"tp {tp_degree} x dp {dp_degree} != {world} ranks"
        self._tp_degree, self._dp_degree = tp_degree, dp_degree
        self._tp_rank_list = self._dp_rank_list = None
        self._tp_process_group = self._dp_process_group = None
        # every rank creates every group (collective requirement of new_group)
        for d in range(dp_degree):
            ranks_tp = [self._rank_list[d * tp_degree + t] for t in range(tp_degree)]
            g = _get_group(ranks_tp, backend)
            if self._rank in ranks_tp:
                self._tp_rank_list, self._tp_process_group = ranks_tp, g
        for t in range(tp_degree):
            ranks_dp = [self._rank_list[d * tp_degree + t] for d in range(dp_degree)]
            g = _get_group(ranks_dp, backend)
            if self._rank in ranks_dp:
                self._dp_rank_list, self._dp_process_group = ranks_dp, g
        self._cpu_tp = self._cpu_dp = None
        self.is_init = True

    def set_cpu_groups(self) -> None:
        if self._cpu_tp is not None:
            return
        for d in range(self._dp_degree):
            ranks_tp = [self._rank_list[d * self._tp_degree + t] for t in range(self._tp_degree)]
            g = _get_group(ranks_tp, "gloo")
            if self._rank in ranks_tp:
                self._cpu_tp = g
        for t in range(self._tp_degree):
            ranks_dp = [self._rank_list[d * self._tp_degree + t] for d in range(self._dp_degree)]
            g = _get_group(ranks_dp, "gloo")
            if self._rank in ranks_dp:
                self._cpu_dp = g

    @property
    def has_cpu_groups(self) -> bool:
        return self._cpu_tp is not None

    def __repr__(self) -> str:
        if not getattr(self, "is_init", False):
            return "ProcessGroup(not initialised)"
        return f"ProcessGroup(ranks={self._rank_list}, rank={self._rank}, dp={self._dp_degree}, tp={self._tp_degree})"

    def __eq__(self, other) -> bool:
        if not isinstance(other, ProcessGroup):
            return False
        keys = ("_rank", "_rank_list", "_tp_degree", "_dp_degree", "_tp_rank_list", "_dp_rank_list")
        return all(getattr(self, k, None) == getattr(other, k, None) for k in keys)

    def __hash__(self) -> int:
        return hash((self._rank, tuple(self._rank_list), self._tp_degree, self._dp_degree))

    def rank(self) -> int:
        return self._rank

    def ranks_in_group(self) -> List[int]:
        return self._rank_list

    def world_size(self) -> int:
        return len(self._rank_list)

    def tp_rank_list(self) -> List[int]:
        return self._tp_rank_list

    def dp_rank_list(self) -> List[int]:
        return self._dp_rank_list

    def tp_local_rank(self) -> int:
        return self._tp_rank_list.index(self._rank)

    def dp_local_rank(self) -> int:
        return self._dp_rank_list.index(self._rank)

    def dp_world_size(self) -> int:
        return self._dp_degree

    def tp_world_size(self) -> int:
        return self._tp_degree

    def dp_process_group(self):
        return self._dp_process_group

    def tp_process_group(self):
        return self._tp_process_group

    def cpu_dp_process_group(self):
        assert self.has_cpu_groups, "call set_cpu_groups() first"
        return self._cpu_dp

    def cpu_tp_process_group(self):
        assert self.has_cpu_groups, "call set_cpu_groups() first"
        return self._cpu_tp

    def get_ranks_in_dp(self) -> List[int]:
        return self._dp_rank_list

    def get_ranks_in_tp(self) -> List[int]:
        return self._tp_rank_list
