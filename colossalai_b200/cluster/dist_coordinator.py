"""Rank helper singleton.  Parity: reference `colossalai/cluster/dist_coordinator.py:11-201`."""
from __future__ import annotations

import functools
import os
from contextlib import contextmanager
from typing import Optional

import torch.distributed as dist
from torch.distributed import ProcessGroup

from ..context import SingletonMeta

__all__ = ["DistCoordinator"]


class DistCoordinator(metaclass=SingletonMeta):
    def __init__(self) -> None:
        assert dist.is_initialized(), "DistCoordinator requires an initialised process group (call launch first)"
        self._rank = dist.get_rank()
        self._world_size = dist.get_world_size()
        self._local_rank = int(os.environ.get("LOCAL_RANK", -1))

    @property
    def rank(self) -> int:
        return self._rank

    @property
    def world_size(self) -> int:
        return self._world_size

    @property
    def local_rank(self) -> int:
        return self._local_rank

    def _assert_local_rank_set(self) -> None:
        assert self._local_rank >= 0, "LOCAL_RANK is not set; launch with torchrun / colossalai_b200 run"

    def is_master(self, process_group: Optional[ProcessGroup] = None) -> bool:
        return dist.get_rank(group=process_group) == 0

    def is_node_master(self) -> bool:
        self._assert_local_rank_set()
        return self._local_rank == 0

    def is_last_process(self, process_group: Optional[ProcessGroup] = None) -> bool:
        return dist.get_rank(group=process_group) == dist.get_world_size(group=process_group) - 1

    def print_on_master(self, msg: str, process_group: Optional[ProcessGroup] = None) -> None:
        if self.is_master(process_group):
            print(msg, flush=True)

    def print_on_node_master(self, msg: str) -> None:
        if self.is_node_master():
            print(msg, flush=True)

    @contextmanager
    def priority_execution(self, executor_rank: int = 0, process_group: Optional[ProcessGroup] = None):
        """`executor_rank` runs the body first, everyone else after it has finished (e.g. dataset download)."""
        is_executor = dist.get_rank(group=process_group) == executor_rank
        if not is_executor:
            self.block_all(process_group)
        yield
        if is_executor:
            self.block_all(process_group)

    def destroy(self, process_group: Optional[ProcessGroup] = None) -> None:
        dist.destroy_process_group(process_group)

    def block_all(self, process_group: Optional[ProcessGroup] = None) -> None:
        dist.barrier(group=process_group)

    def on_master_only(self, process_group: Optional[ProcessGroup] = None):
        def decorator(func):
            @functools.wraps(func)
            def wrapper(*a, **k):
                if self.is_master(process_group):
                    return func(*a, **k)

            return wrapper

        return decorator
