"""RNG discipline + sequence-parallel gradient helpers + ring communicator.

Parity: reference `colossalai/shardformer/layer/utils.py:52-330` (`SeqParallelUtils`, `Randomizer`),
`:475` (`RingComm`).
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.distributed import ProcessGroup

from ...parallel import comm


class SeqParallelUtils:
    """Norm/bias params that only see a sequence slice under split_gather/ring SP have PARTIAL grads:
    mark them, then all-reduce their grads over the TP group after backward."""

    @staticmethod
    def marked_as_sp_partial_derived_param(param: torch.Tensor) -> None:
        setattr(param, "partial_derived", True)

    @staticmethod
    def is_sp_partial_derived_param(param: torch.Tensor) -> bool:
        return getattr(param, "partial_derived", False)

    @staticmethod
    def allreduce_partial_data_grad(process_group: ProcessGroup, model: Optional[nn.Module] = None,
                                    grads: Optional[List[torch.Tensor]] = None) -> None:
        assert (model is None) != (grads is None), "pass exactly one of model / grads"
        if comm.group_size(process_group) == 1:
            return
        if model is not None:
            grads = [p.grad for p in model.parameters()
                     if p.grad is not None and SeqParallelUtils.is_sp_partial_derived_param(p)]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1).float() for g in grads])
        dist.all_reduce(flat, group=process_group)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


class Randomizer:
    """Holds a private RNG state (cpu + device) that can be swapped in with `fork_rng`, so TP-sharded weights /
    dropout on sharded activations use a per-rank stream while replicated tensors keep the global stream."""

    _INDEX = 0

    def __init__(self, seed: int) -> None:
        self.seed = seed
        cpu_state = torch.get_rng_state()
        torch.manual_seed(seed)
        self.cpu_rng_state = torch.get_rng_state()
        torch.set_rng_state(cpu_state)
        self.device_rng_state = None
        if torch.cuda.is_available():
            dev_state = torch.cuda.get_rng_state()
            torch.cuda.manual_seed(seed)
            self.device_rng_state = torch.cuda.get_rng_state()
            torch.cuda.set_rng_state(dev_state)

    @contextmanager
    def fork_rng(self, enable_cpu: bool = False):
        dev_backup = torch.cuda.get_rng_state() if self.device_rng_state is not None else None
        cpu_backup = torch.get_rng_state() if enable_cpu else None
        try:
            if dev_backup is not None:
                torch.cuda.set_rng_state(self.device_rng_state)
            if enable_cpu:
                torch.set_rng_state(self.cpu_rng_state)
            yield
        finally:
            if dev_backup is not None:
                self.device_rng_state = torch.cuda.get_rng_state()
                torch.cuda.set_rng_state(dev_backup)
            if enable_cpu:
                self.cpu_rng_state = torch.get_rng_state()
                torch.set_rng_state(cpu_backup)

    @staticmethod
    def index() -> int:
        return Randomizer._INDEX

    @staticmethod
    def increment_index() -> None:
        Randomizer._INDEX += 1

    @staticmethod
    def reset_index() -> None:
        Randomizer._INDEX = 0

    @staticmethod
    def is_randomizer_index_synchronized(process_group: Optional[ProcessGroup] = None) -> bool:
        if not dist.is_initialized():
            return True
        idx = torch.tensor([Randomizer._INDEX], dtype=torch.int64)
        if dist.get_backend(process_group) == "nccl":
            idx = idx.cuda()
        lst = [torch.zeros_like(idx) for _ in range(dist.get_world_size(process_group))]
        dist.all_gather(lst, idx, group=process_group)
        return all(int(x) == int(lst[0]) for x in lst)

    @staticmethod
    def synchronize_index(process_group: Optional[ProcessGroup] = None) -> None:
        if not dist.is_initialized():
            return
        idx = torch.tensor([Randomizer._INDEX], dtype=torch.int64)
        if dist.get_backend(process_group) == "nccl":
            idx = idx.cuda()
        src = dist.get_process_group_ranks(process_group)[0] if process_group is not None else 0
        dist.broadcast(idx, src=src, group=process_group)
        Randomizer._INDEX = int(idx.item())


def create_randomizer_with_offset(seed: int, process_group: Optional[ProcessGroup] = None,
                                  offset_by_rank: bool = True, offset_by_index: bool = True) -> Randomizer:
    """seed + rank-in-group (different shards initialise differently) + a per-layer index."""
    if offset_by_index:
        seed += Randomizer.index()
        Randomizer.increment_index()
    if offset_by_rank and dist.is_initialized():
        seed += 1000003 * (comm.group_rank(process_group) + 1) if comm.group_size(process_group) > 1 else 0
    return Randomizer(seed % (2**31))


class RingComm:
    """Next/prev P2P exchange along a process group (ring attention KV / dKV circulation on the NCCL backend)."""

    def __init__(self, process_group: Optional[ProcessGroup]) -> None:
        self.group = process_group
        self.rank = comm.group_rank(process_group)
        self.world_size = comm.group_size(process_group)
        self._handles = []
        ranks = dist.get_process_group_ranks(process_group) if (dist.is_initialized() and process_group is not None) \
            else list(range(self.world_size))
        self.send_rank = ranks[(self.rank + 1) % self.world_size]
        self.recv_rank = ranks[(self.rank - 1) % self.world_size]

    def send_recv(self, send_tensor: torch.Tensor, recv_tensor: Optional[torch.Tensor] = None,
                  commit: bool = True) -> torch.Tensor:
        if recv_tensor is None:
            recv_tensor = torch.empty_like(send_tensor)
        ops = [dist.P2POp(dist.isend, send_tensor, self.send_rank, self.group),
               dist.P2POp(dist.irecv, recv_tensor, self.recv_rank, self.group)]
        if self.rank % 2 == 1:
            ops.reverse()
        self._pending = getattr(self, "_pending", []) + ops
        if commit:
            self.commit()
        return recv_tensor

    def commit(self) -> None:
        ops, self._pending = getattr(self, "_pending", []), []
        if ops:
            self._handles = dist.batch_isend_irecv(ops)

    def wait(self) -> None:
        for h in self._handles:
            h.wait()
        self._handles = []
