"""ChunkManager: registry of chunks per group, tensor -> chunk lookup, access / release / move / reduce.
Parity: reference `colossalai/zero/gemini/chunk/manager.py:14-317`."""
from __future__ import annotations

from collections import deque
from typing import Deque, Dict, Iterable, List, Optional, Set, Tuple

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup

from ....accelerator import get_accelerator
from .chunk import Chunk, ChunkFullError, TensorState

__all__ = ["ChunkManager"]


class ChunkManager:
    def __init__(self, chunk_configuration: Dict[int, Dict], init_device: Optional[torch.device] = None,
                 reuse_fp16_chunk: bool = True, max_prefetch: int = 0) -> None:
        self.device = init_device or get_accelerator().get_current_device()
        self.dp_degree_chunk_size_dict: Dict[int, int] = {}
        self.kwargs_config = chunk_configuration
        for k, v in self.kwargs_config.items():
            self.dp_degree_chunk_size_dict[k] = v.pop("chunk_size")
            v["init_device"] = self.device
        self.chunk_groups: Dict[str, Deque[Chunk]] = {}
        self.tensor_chunk_map: Dict[torch.Tensor, Chunk] = {}
        self.accessed_chunks: Set[Chunk] = set()
        self.accessed_mem: int = 0
        self.total_mem: Dict[str, int] = {"cpu": 0, "cuda": 0}
        self.reuse_fp16_chunk = reuse_fp16_chunk
        self.overflow_counter = torch.zeros(1, dtype=torch.int32, device=get_accelerator().get_current_device())
        self._prefetch_stream = get_accelerator().Stream() if max_prefetch else None
        self._async_works: Dict[Chunk, object] = {}

    def register_tensor(self, tensor: torch.Tensor, group_type: str, config_key: int, zero_group: ProcessGroup,
                        extra_dp_group: ProcessGroup = None, cpu_offload: bool = False,
                        pin_memory: bool = False) -> None:
        assert tensor not in self.tensor_chunk_map
        assert isinstance(tensor, torch.Tensor)
        assert config_key in self.dp_degree_chunk_size_dict
        chunk_size = self.dp_degree_chunk_size_dict[config_key]
        chunk_kwargs = self.kwargs_config[config_key]
        group_name = f"{group_type}_{config_key}"
        chunk_group = self.__get_chunk_group(group_name)
        try:
            chunk_group[-1].append_tensor(tensor)
        except (IndexError, ChunkFullError):
            if chunk_group:
                self.__close_one_chunk(chunk_group[-1])
            if tensor.numel() > chunk_size:
                ws = dist.get_world_size(zero_group) if dist.is_initialized() else 1
                chunk_size = tensor.numel() + ((ws - (tensor.numel() % ws)) % ws)
            chunk = Chunk(chunk_size=chunk_size, zero_group=zero_group, dtype=tensor.dtype,
                          cpu_shard_init=cpu_offload, pin_memory=pin_memory, extra_dp_group=extra_dp_group,
                          **chunk_kwargs)
            chunk_group.append(chunk)
            chunk.append_tensor(tensor)
            self.__add_memory_usage(chunk.memory_usage)
        self.tensor_chunk_map[tensor] = chunk_group[-1]

    def close_all_groups(self) -> None:
        for name in self.chunk_groups:
            self.__close_one_chunk(self.chunk_groups[name][-1])

    def access_chunk(self, chunk: Chunk, async_access: bool = False):
        if chunk in self.accessed_chunks:
            return None
        self.__sub_memory_usage(chunk.memory_usage)
        if chunk.device_type == "cpu" and not chunk.is_gathered:
            chunk.shard_move(get_accelerator().get_current_device())
        work = self.__add_accessed_chunk(chunk, async_access=async_access)
        self.__add_memory_usage(chunk.memory_usage)
        return work

    def release_chunk(self, chunk: Chunk) -> None:
        if chunk not in self.accessed_chunks:
            return
        if chunk.can_release:
            self.__sub_memory_usage(chunk.memory_usage)
            self.__sub_accessed_chunk(chunk)
            self.__add_memory_usage(chunk.memory_usage)

    def move_chunk(self, chunk: Chunk, device: torch.device, force_copy: bool = False, async_move: bool = False) -> None:
        if not chunk.can_move or chunk.device_type == torch.device(device).type:
            return
        self.__sub_memory_usage(chunk.memory_usage)
        chunk.shard_move(device, force_copy, non_blocking=async_move)
        self.__add_memory_usage(chunk.memory_usage)

    def trans_tensor_state(self, tensor: torch.Tensor, state: TensorState) -> None:
        self.tensor_chunk_map[tensor].tensor_trans_state(tensor, state)

    def reduce_chunk(self, chunk: Chunk, async_op: bool = False) -> bool:
        if not chunk.can_reduce:
            return False
        self.__sub_memory_usage(chunk.memory_usage)
        chunk.reduce(async_op=async_op)
        self.__sub_accessed_chunk(chunk)
        self.__add_memory_usage(chunk.memory_usage)
        return True

    def fake_release_chunk(self, chunk: Chunk) -> None:
        assert chunk.keep_gathered and chunk.pg_size == 1
        self.__sub_accessed_chunk(chunk)

    def copy_tensor_to_chunk_slice(self, tensor: torch.Tensor, data: torch.Tensor) -> None:
        self.tensor_chunk_map[tensor].copy_tensor_to_chunk_slice(tensor, data)

    def get_chunk(self, tensor: torch.Tensor) -> Chunk:
        return self.tensor_chunk_map[tensor]

    def get_cuda_movable_chunks(self) -> List[Chunk]:
        return [c for c in self.accessed_chunks if c.can_release]

    def get_chunks(self, tensors: Iterable[torch.Tensor]) -> Tuple[Chunk, ...]:
        out = {}
        for t in tensors:
            out[self.get_chunk(t)] = None
        return tuple(out.keys())

    def add_extern_static_tensor(self, tensor: torch.Tensor) -> None:
        self.total_mem[tensor.device.type if tensor.device.type == "cpu" else "cuda"] += tensor.numel() * tensor.element_size()

    def init_grad_chunk(self, chunk: Chunk) -> Chunk:
        if chunk.grad_chunk is not None:
            self.__sub_memory_usage(chunk.grad_chunk.memory_usage)
        g = chunk.init_grad_chunk()
        self.__add_memory_usage(g.memory_usage)
        self.__add_accessed_chunk_no_gather(g)
        return g

    def rearrange_accumulated_grad_chunk(self, chunk: Chunk) -> Chunk:
        return chunk.grad_chunk

    def all_chunks(self) -> List[Chunk]:
        return [c for g in self.chunk_groups.values() for c in g]

    def __repr__(self) -> str:
        lines = ["Chunk Manager Information:", f"Total memory: " + ", ".join(f"{k}={v}" for k, v in self.total_mem.items())]
        for name, group in self.chunk_groups.items():
            lines.append(f"Group {name}:")
            for i, c in enumerate(group):
                lines.append(f"[{i}] {c}")
        return "\n".join(lines)

    # ------------------------------------------------------------------ internals
    def __get_chunk_group(self, group_name: str) -> Deque[Chunk]:
        if group_name not in self.chunk_groups:
            self.chunk_groups[group_name] = deque()
        return self.chunk_groups[group_name]

    def __close_one_chunk(self, chunk: Chunk) -> None:
        if chunk.chunk_temp is None:
            return
        self.__sub_memory_usage(chunk.memory_usage)
        chunk.close_chunk()
        self.__add_memory_usage(chunk.memory_usage)
        if chunk.is_gathered:
            self.accessed_chunks.add(chunk)

    def __sub_memory_usage(self, usage: Dict[str, int]) -> None:
        for k, v in usage.items():
            self.total_mem[k] -= v

    def __add_memory_usage(self, usage: Dict[str, int]) -> None:
        for k, v in usage.items():
            self.total_mem[k] += v

    def __add_accessed_chunk(self, chunk: Chunk, async_access: bool = False):
        work = chunk.access_chunk(async_access=async_access)
        self.accessed_chunks.add(chunk)
        self.accessed_mem += chunk.chunk_mem
        return work

    def __add_accessed_chunk_no_gather(self, chunk: Chunk) -> None:
        self.accessed_chunks.add(chunk)
        self.accessed_mem += chunk.chunk_mem

    def __sub_accessed_chunk(self, chunk: Chunk) -> None:
        chunk.release_chunk()
        self.accessed_chunks.discard(chunk)
        self.accessed_mem -= chunk.chunk_mem
