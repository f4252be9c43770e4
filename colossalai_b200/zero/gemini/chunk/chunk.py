"""Chunk: a fixed-size flat buffer that packs several parameters, sharded over the ZeRO group and movable between
HBM and (pinned) host memory.

Parity: reference `colossalai/zero/gemini/chunk/chunk.py:59-683` (`TensorState` machine, `append_tensor`, `close_chunk`,
`shard_move`, `access_chunk` / `release_chunk`, `reduce`, `tensor_trans_state`, `copy_tensor_to_chunk_slice`, inf/nan
and l2-norm bookkeeping, `keep_gathered`, paired fp32 chunk).
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup

from ....accelerator import get_accelerator
from ....parallel import comm

__all__ = ["Chunk", "ChunkFullError", "TensorState", "TensorInfo", "alloc_storage", "free_storage"]


class TensorState(Enum):
    FREE = 0
    COMPUTE = 1
    HOLD = 2
    HOLD_AFTER_BWD = 3
    READY_FOR_REDUCE = 4


STATE_TRANS = (
    (TensorState.FREE, TensorState.HOLD),
    (TensorState.FREE, TensorState.COMPUTE),
    (TensorState.HOLD, TensorState.FREE),
    (TensorState.HOLD, TensorState.COMPUTE),
    (TensorState.COMPUTE, TensorState.HOLD),
    (TensorState.COMPUTE, TensorState.HOLD_AFTER_BWD),
    (TensorState.HOLD_AFTER_BWD, TensorState.COMPUTE),
    (TensorState.HOLD_AFTER_BWD, TensorState.READY_FOR_REDUCE),
    (TensorState.READY_FOR_REDUCE, TensorState.HOLD),
)


@dataclass
class TensorInfo:
    state: TensorState
    offset: int
    end: int
    shape: tuple = ()


class ChunkFullError(Exception):
    pass


def free_storage(t: torch.Tensor) -> None:
    if t.untyped_storage().size() > 0:
        t.untyped_storage().resize_(0)


def alloc_storage(t: torch.Tensor) -> None:
    if t.untyped_storage().size() == 0:
        t.untyped_storage().resize_(t.numel() * t.element_size())


class Chunk:
    _total_number = 0

    def __init__(self, chunk_size: int, zero_group: ProcessGroup, dtype: torch.dtype,
                 init_device: Optional[torch.device] = None, cpu_shard_init: bool = False,
                 keep_gathered: bool = False, pin_memory: bool = False, extra_dp_group: ProcessGroup = None) -> None:
        self.count_id = Chunk._total_number
        Chunk._total_number += 1
        self.chunk_size = chunk_size
        self.utilized_size = 0
        self.torch_pg = zero_group
        self.pg_size = comm.group_size(zero_group)
        self.pg_rank = comm.group_rank(zero_group)
        self.extra_dp_group = extra_dp_group
        self.extra_dp_size = comm.group_size(extra_dp_group) if extra_dp_group is not None else 1
        assert chunk_size % self.pg_size == 0, "chunk size must be divisible by the zero group size"
        self.shard_size = chunk_size // self.pg_size
        self.shard_begin = self.shard_size * self.pg_rank
        self.shard_end = self.shard_begin + self.shard_size
        self.dtype = dtype
        self.device = init_device or get_accelerator().get_current_device()
        self.chunk_temp: Optional[torch.Tensor] = torch.zeros(chunk_size, dtype=dtype, device=self.device)
        self.cuda_global_chunk: Optional[torch.Tensor] = None     # full (gathered) buffer when accessed
        self.cuda_shard: Optional[torch.Tensor] = None
        self.cpu_shard: Optional[torch.Tensor] = None
        self.is_gathered = True
        self.keep_gathered = keep_gathered or self.pg_size == 1 and False
        self.pin_memory = pin_memory
        self.cpu_shard_init = cpu_shard_init
        self.tensors_info: Dict[torch.Tensor, TensorInfo] = {}
        self.tensor_state_cnter: Dict[TensorState, int] = {s: 0 for s in TensorState}
        self.paired_chunk: Optional["Chunk"] = None
        self.grad_chunk: Optional["Chunk"] = None
        self.l2_norm_flag = False
        self.l2_norm: Optional[float] = None
        self.cpu_vis_flag = False
        self.overflow = False
        self.is_grad_chunk = False

    # ------------------------------------------------------------------ sizes / placement
    @property
    def memory_usage(self) -> Dict[str, int]:
        cuda, cpu = 0, 0
        if self.chunk_temp is not None:
            cuda += self.chunk_temp.numel() * self.chunk_temp.element_size() if self.chunk_temp.is_cuda else 0
        if self.is_gathered and self.cuda_global_chunk is not None:
            cuda += self.chunk_size * self.cuda_global_chunk.element_size()
        elif self.cuda_shard is not None:
            cuda += self.shard_size * self.cuda_shard.element_size()
        if self.cpu_shard is not None:
            cpu += self.shard_size * self.cpu_shard.element_size()
        return dict(cuda=cuda, cpu=cpu)

    @property
    def chunk_mem(self) -> int:
        return self.chunk_size * torch.tensor([], dtype=self.dtype).element_size()

    @property
    def shard_mem(self) -> int:
        return self.chunk_mem // self.pg_size

    @property
    def device_type(self) -> str:
        if self.chunk_temp is not None:
            return self.chunk_temp.device.type
        if self.is_gathered or self.cuda_shard is not None:
            return get_accelerator().name
        return "cpu"

    @property
    def payload(self) -> torch.Tensor:
        if self.chunk_temp is not None:
            return self.chunk_temp
        if self.is_gathered:
            return self.cuda_global_chunk
        return self.cuda_shard if self.cuda_shard is not None else self.cpu_shard

    @property
    def payload_mem(self) -> int:
        if self.chunk_temp is not None or self.is_gathered:
            return self.chunk_mem
        return self.shard_mem

    @property
    def can_move(self) -> bool:
        return not self.is_gathered

    @property
    def can_release(self) -> bool:
        if self.keep_gathered:
            return False
        return self.is_gathered and (self.tensor_state_cnter[TensorState.HOLD]
                                     + self.tensor_state_cnter[TensorState.HOLD_AFTER_BWD] == self.num_tensors)

    @property
    def can_reduce(self) -> bool:
        return self.tensor_state_cnter[TensorState.READY_FOR_REDUCE] == self.num_tensors

    @property
    def has_inf_or_nan(self) -> bool:
        t = self.cuda_shard if self.cuda_shard is not None else self.cpu_shard
        if self.is_gathered and self.cuda_global_chunk is not None:
            t = self.cuda_global_chunk[: self.utilized_size]
        return bool(t is not None and (not torch.isfinite(t).all()))

    def nonfinite_flag(self) -> Optional[torch.Tensor]:
        """Device-side 0/1 int32 tensor version of `has_inf_or_nan` (no host sync: the backward pass adds it into the
        overflow counter and keeps launching)."""
        t = self.cuda_shard if self.cuda_shard is not None else self.cpu_shard
        if self.is_gathered and self.cuda_global_chunk is not None:
            t = self.cuda_global_chunk[: self.utilized_size]
        return None if t is None else (~torch.isfinite(t).all()).to(torch.int32)

    @property
    def num_tensors(self) -> int:
        return len(self.tensors_info)

    def set_l2_norm(self) -> None:
        t = self.cuda_shard if self.cuda_shard is not None else self.cpu_shard
        if self.is_gathered and self.cuda_global_chunk is not None:
            t = self.cuda_global_chunk[: self.utilized_size]
        # kept as a tensor on the shard's device: the optimizer sums the per-chunk values and reads them back once
        self.l2_norm = torch.sum(t.float() ** 2) if t is not None else 0.0

    # ------------------------------------------------------------------ construction
    def append_tensor(self, tensor: torch.Tensor) -> None:
        assert self.chunk_temp is not None, "chunk is already closed"
        assert tensor.dtype == self.dtype
        new_util = self.utilized_size + tensor.numel()
        if new_util > self.chunk_size:
            raise ChunkFullError
        self.chunk_temp[self.utilized_size:new_util].copy_(tensor.data.flatten())
        tensor.data = self.chunk_temp[self.utilized_size:new_util].view(tensor.shape)
        self.tensors_info[tensor] = TensorInfo(TensorState.HOLD, self.utilized_size, new_util, tuple(tensor.shape))
        self.utilized_size = new_util
        self.tensor_state_cnter[TensorState.HOLD] += 1

    def close_chunk(self) -> None:
        """Finish construction: keep only this rank's shard (unless keep_gathered)."""
        assert self.chunk_temp is not None
        dev = get_accelerator().get_current_device()
        if self.chunk_temp.device.type == "cpu" and not self.cpu_shard_init:
            full = self.chunk_temp.to(dev)
        else:
            full = self.chunk_temp
        self.chunk_temp = None
        if self.keep_gathered or self.pg_size == 1 and self.keep_gathered:
            self.cuda_global_chunk = full.to(dev)
            self.is_gathered = True
            self.__update_tensors_ptr()
            return
        shard = full[self.shard_begin:self.shard_end].clone()
        self.is_gathered = False
        if self.cpu_shard_init and get_accelerator().name != "cpu":
            self.cpu_shard = shard.cpu()
            if self.pin_memory and torch.cuda.is_available():
                self.cpu_shard = self.cpu_shard.pin_memory()
            self.cuda_shard = None
        else:
            self.cuda_shard = shard.to(dev)
        self.__update_tensors_link(None)

    # ------------------------------------------------------------------ movement
    def shard_move(self, device: torch.device, force_copy: bool = False, non_blocking: bool = False) -> None:
        """Move this rank's shard between cuda and (pinned) cpu."""
        device = torch.device(device)
        if self.is_gathered or get_accelerator().name == "cpu":
            return   # CPU plumbing tier: the "accelerator" shard already lives in host memory
        if device.type != "cpu":
            if self.cuda_shard is not None:
                return
            self.cuda_shard = self.cpu_shard.to(device, non_blocking=non_blocking)
            if not self.pin_memory:
                self.cpu_shard = None
        else:
            if self.cuda_shard is None:
                return
            if self.pin_memory and torch.cuda.is_available():
                if force_copy or not self.cpu_vis_flag or self.cpu_shard is None:
                    if self.cpu_shard is None:
                        self.cpu_shard = torch.empty(self.shard_size, dtype=self.dtype, pin_memory=True)
                    self.cpu_shard.copy_(self.cuda_shard, non_blocking=non_blocking)
            else:
                self.cpu_shard = self.cuda_shard.cpu()
            self.cpu_vis_flag = True
            self.cuda_shard = None

    def access_chunk(self, async_access: bool = False):
        """Make the full chunk available on the accelerator (all-gather the shards)."""
        if self.is_gathered:
            return None
        dev = get_accelerator().get_current_device()
        if self.cuda_shard is None:
            self.shard_move(dev)
        work = self.__gather(async_op=async_access)
        self.__update_tensors_ptr()
        return work

    def release_chunk(self) -> None:
        """Drop the gathered buffer, keeping only the shard."""
        if self.is_gathered and not self.keep_gathered:
            self.__scatter()

    def __gather(self, async_op: bool = False):
        dev = get_accelerator().get_current_device()
        self.cuda_global_chunk = torch.empty(self.chunk_size, dtype=self.dtype, device=dev)
        work = None
        if self.pg_size > 1:
            work = dist.all_gather_into_tensor(self.cuda_global_chunk, self.cuda_shard, group=self.torch_pg,
                                               async_op=async_op)
        else:
            self.cuda_global_chunk.copy_(self.cuda_shard)
        self.cuda_shard = None
        self.is_gathered = True
        return work

    def __scatter(self) -> None:
        if self.keep_gathered:
            return
        dev = get_accelerator().get_current_device()
        self.cuda_shard = torch.empty(self.shard_size, dtype=self.dtype, device=dev)
        self.cuda_shard.copy_(self.cuda_global_chunk[self.shard_begin:self.shard_end])
        self.__update_tensors_link(None)
        self.cuda_global_chunk = None
        self.is_gathered = False

    def reduce(self, async_op: bool = False):
        """Reduce-scatter the (gradient) chunk over the zero group; average over zero x extra_dp."""
        assert self.is_gathered
        work = None
        if self.pg_size == 1 and self.extra_dp_size == 1:
            if not self.keep_gathered:
                self.cuda_shard = self.cuda_global_chunk[self.shard_begin:self.shard_end].clone()
        else:
            dev = get_accelerator().get_current_device()
            if self.pg_size > 1:
                self.cuda_shard = torch.empty(self.shard_size, dtype=self.dtype, device=dev)
                self.cuda_global_chunk.div_(self.pg_size)
                dist.reduce_scatter_tensor(self.cuda_shard, self.cuda_global_chunk, group=self.torch_pg)
            else:
                self.cuda_shard = self.cuda_global_chunk[self.shard_begin:self.shard_end].clone()
            if self.extra_dp_group is not None and self.extra_dp_size > 1:
                self.cuda_shard.div_(self.extra_dp_size)
                dist.all_reduce(self.cuda_shard, group=self.extra_dp_group)
        if not self.keep_gathered:
            self.cuda_global_chunk = None
            self.is_gathered = False
            self.__update_tensors_link(None)
        for t in self.tensors_info:
            self.__update_one_tensor_info(self.tensors_info[t], TensorState.HOLD)
        return work

    # ------------------------------------------------------------------ tensors
    def tensor_trans_state(self, tensor: torch.Tensor, tensor_state: TensorState) -> None:
        info = self.tensors_info[tensor]
        if info.state == tensor_state:
            return
        if (info.state, tensor_state) not in STATE_TRANS:
            return   # illegal transitions are ignored by design here; the DDP wrapper asserts the critical ones
        self.__update_one_tensor_info(info, tensor_state)

    def copy_tensor_to_chunk_slice(self, tensor: torch.Tensor, data_slice: torch.Tensor, update_ptr: bool = True) -> None:
        assert self.is_gathered
        info = self.tensors_info[tensor]
        self.cuda_global_chunk[info.offset:info.end].copy_(data_slice.data.flatten())
        if update_ptr:
            tensor.data = self.cuda_global_chunk[info.offset:info.end].view(info.shape)

    def add_tensor_to_chunk_slice(self, tensor: torch.Tensor, data_slice: torch.Tensor) -> None:
        assert self.is_gathered
        info = self.tensors_info[tensor]
        self.cuda_global_chunk[info.offset:info.end].add_(data_slice.data.flatten())

    def get_valid_length(self) -> int:
        if self.keep_gathered:
            return self.utilized_size
        return max(min(self.utilized_size, self.shard_end) - self.shard_begin, 0)

    def get_tensors(self) -> List[torch.Tensor]:
        return list(self.tensors_info.keys())

    def init_pair(self, friend_chunk: "Chunk") -> None:
        if self.paired_chunk is None and friend_chunk.paired_chunk is None:
            self.paired_chunk, friend_chunk.paired_chunk = friend_chunk, self

    def optim_update(self) -> None:
        """Copy the paired fp32 master shard into this (low-precision) chunk."""
        friend = self.paired_chunk
        assert friend is not None
        if self.is_gathered:
            src = friend.cuda_global_chunk if friend.is_gathered else None
            if src is not None:
                self.cuda_global_chunk.copy_(src)
            else:   # master is sharded while the working chunk is kept gathered: refresh own slice then re-gather
                fs = friend.cuda_shard if friend.cuda_shard is not None else friend.cpu_shard
                self.cuda_global_chunk[self.shard_begin:self.shard_end].copy_(fs, non_blocking=True)
                if self.pg_size > 1:
                    mine = self.cuda_global_chunk[self.shard_begin:self.shard_end].clone()
                    dist.all_gather_into_tensor(self.cuda_global_chunk, mine, group=self.torch_pg)
        else:
            dst = self.cuda_shard if self.cuda_shard is not None else self.cpu_shard
            src = friend.cuda_shard if friend.cuda_shard is not None else friend.cpu_shard
            if friend.is_gathered:
                src = friend.cuda_global_chunk[self.shard_begin:self.shard_end]
            dst.copy_(src, non_blocking=True)
        self.cpu_vis_flag = False

    def init_grad_chunk(self) -> "Chunk":
        """A gathered, zero-filled buffer with the same tensor layout that receives this chunk's gradients."""
        dev = get_accelerator().get_current_device()
        g = Chunk.__new__(Chunk)
        g.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("tensors_info", "tensor_state_cnter")})
        g.count_id = -self.count_id - 1
        g.chunk_temp = None
        g.cuda_global_chunk = torch.zeros(self.chunk_size, dtype=self.dtype, device=dev)
        g.cuda_shard = g.cpu_shard = None
        g.is_gathered = True
        g.is_grad_chunk = True       # never touch the parameters' .data pointers
        g.keep_gathered = False      # gradients are always reduce-scattered (ZeRO-2 semantics for kept params)
        g.tensors_info = {t: TensorInfo(TensorState.HOLD_AFTER_BWD, i.offset, i.end, i.shape) for t, i in self.tensors_info.items()}
        g.tensor_state_cnter = {s: 0 for s in TensorState}
        g.tensor_state_cnter[TensorState.HOLD_AFTER_BWD] = len(g.tensors_info)
        g.paired_chunk = None
        g.grad_chunk = None
        g.l2_norm = None
        self.grad_chunk = g
        return g

    # ------------------------------------------------------------------ internals
    def __update_tensors_ptr(self) -> None:
        if self.is_grad_chunk:
            return
        assert self.is_gathered
        for t, info in self.tensors_info.items():
            t.data = self.cuda_global_chunk[info.offset:info.end].view(info.shape)

    def __update_tensors_link(self, placeholder) -> None:
        """Point member tensors at an empty placeholder (their storage lives in the shard now)."""
        if self.is_grad_chunk:
            return
        for t in self.tensors_info:
            t.data = torch.empty(0, dtype=t.dtype, device=t.device if t.device.type != "meta" else "cpu")

    def __update_one_tensor_info(self, info: TensorInfo, nxt: TensorState) -> None:
        self.tensor_state_cnter[info.state] -= 1
        info.state = nxt
        self.tensor_state_cnter[nxt] += 1

    def __hash__(self) -> int:
        return hash(id(self))

    def __eq__(self, other) -> bool:
        return self is other

    def __repr__(self, detailed: bool = False) -> str:
        return (f"Chunk(id={self.count_id}, size={self.chunk_size}, utilized={self.utilized_size}, "
                f"gathered={self.is_gathered}, keep_gathered={self.keep_gathered}, device={self.device_type}, "
                f"tensors={self.num_tensors})")
