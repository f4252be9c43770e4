"""ChunkManager: registry of chunks per group, tensor -> chunk lookup, access / release / move / reduce.
Parity: reference `colossalai/zero/gemini/chunk/manager.py:14-317`."""
from __future__ import annotations

import contextlib
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

    # ------------------------------------------------------------------ bookkeeping
    @contextlib.contextmanager
    def _accounted(self, chunk: Chunk):
        """Keep `total_mem` consistent across an operation that changes where a chunk's bytes live: the chunk's footprint
        is taken out of the totals, the operation runs, the new footprint is put back."""
        for dev, n in chunk.memory_usage.items():
            self.total_mem[dev] -= n
        try:
            yield chunk
        finally:
            for dev, n in chunk.memory_usage.items():
                self.total_mem[dev] += n

    def _mark_resident(self, chunk: Chunk) -> None:
        self.accessed_chunks.add(chunk)
        self.accessed_mem += chunk.chunk_mem

    def _drop_resident(self, chunk: Chunk) -> None:
        chunk.release_chunk()
        self.accessed_chunks.discard(chunk)
        self.accessed_mem -= chunk.chunk_mem

    def _seal(self, chunk: Chunk) -> None:
        """Finish filling a chunk (its temporary full-size buffer becomes shard / gathered storage)."""
        if chunk.chunk_temp is None:
            return
        with self._accounted(chunk):
            chunk.close_chunk()
        if chunk.is_gathered:
            self.accessed_chunks.add(chunk)

    # ------------------------------------------------------------------ construction
    def register_tensor(self, tensor: torch.Tensor, group_type: str, config_key: int, zero_group: ProcessGroup,
                        extra_dp_group: ProcessGroup = None, cpu_offload: bool = False,
                        pin_memory: bool = False) -> None:
        """Place `tensor` into the open chunk of its (kind, dp-degree) group, opening a new chunk when it does not fit
        (a tensor larger than the configured chunk size gets a chunk of its own, padded to the group size)."""
        assert isinstance(tensor, torch.Tensor) and tensor not in self.tensor_chunk_map
        assert config_key in self.dp_degree_chunk_size_dict, f"no chunk configuration for dp degree {config_key}"
        group = self.chunk_groups.setdefault(f"{group_type}_{config_key}", deque())
        open_chunk = group[-1] if group else None
        if open_chunk is not None:
            try:
                open_chunk.append_tensor(tensor)
                self.tensor_chunk_map[tensor] = open_chunk
                return
            except ChunkFullError:
                self._seal(open_chunk)
        size = self.dp_degree_chunk_size_dict[config_key]
        if tensor.numel() > size:
            ws = dist.get_world_size(zero_group) if dist.is_initialized() else 1
            size = -(-tensor.numel() // ws) * ws
        chunk = Chunk(chunk_size=size, zero_group=zero_group, dtype=tensor.dtype, cpu_shard_init=cpu_offload,
                      pin_memory=pin_memory, extra_dp_group=extra_dp_group, **self.kwargs_config[config_key])
        chunk.append_tensor(tensor)
        group.append(chunk)
        for dev, n in chunk.memory_usage.items():
            self.total_mem[dev] += n
        self.tensor_chunk_map[tensor] = chunk

    def close_all_groups(self) -> None:
        for group in self.chunk_groups.values():
            self._seal(group[-1])

    # ------------------------------------------------------------------ residency
    def access_chunk(self, chunk: Chunk, async_access: bool = False):
        """Make the full chunk resident on the accelerator (all-gather of the shards); returns the async work, if any."""
        if chunk in self.accessed_chunks:
            return None
        with self._accounted(chunk):
            if chunk.device_type == "cpu" and not chunk.is_gathered:
                chunk.shard_move(get_accelerator().get_current_device())
            work = chunk.access_chunk(async_access=async_access)
            self._mark_resident(chunk)
        return work

    def release_chunk(self, chunk: Chunk) -> None:
        """Back to shards, when no tensor of the chunk is in use."""
        if chunk in self.accessed_chunks and chunk.can_release:
            with self._accounted(chunk):
                self._drop_resident(chunk)

    def move_chunk(self, chunk: Chunk, device: torch.device, force_copy: bool = False, async_move: bool = False) -> None:
        if chunk.can_move and chunk.device_type != torch.device(device).type:
            with self._accounted(chunk):
                chunk.shard_move(device, force_copy, non_blocking=async_move)

    def reduce_chunk(self, chunk: Chunk, async_op: bool = False) -> bool:
        """Reduce-scatter a gradient chunk whose tensors are all ready; the gathered buffer is dropped."""
        if not chunk.can_reduce:
            return False
        with self._accounted(chunk):
            chunk.reduce(async_op=async_op)
            self._drop_resident(chunk)
        return True

    def fake_release_chunk(self, chunk: Chunk) -> None:
        assert chunk.keep_gathered and chunk.pg_size == 1
        self._drop_resident(chunk)

    def init_grad_chunk(self, chunk: Chunk) -> Chunk:
        old = chunk.grad_chunk
        if old is not None:
            for dev, n in old.memory_usage.items():
                self.total_mem[dev] -= n
        g = chunk.init_grad_chunk()
        for dev, n in g.memory_usage.items():
            self.total_mem[dev] += n
        self._mark_resident(g)
        return g

    def rearrange_accumulated_grad_chunk(self, chunk: Chunk) -> Chunk:
        return chunk.grad_chunk

    # ------------------------------------------------------------------ lookups
    def trans_tensor_state(self, tensor: torch.Tensor, state: TensorState) -> None:
        self.tensor_chunk_map[tensor].tensor_trans_state(tensor, state)

    def copy_tensor_to_chunk_slice(self, tensor: torch.Tensor, data: torch.Tensor) -> None:
        self.tensor_chunk_map[tensor].copy_tensor_to_chunk_slice(tensor, data)

    def get_chunk(self, tensor: torch.Tensor) -> Chunk:
        return self.tensor_chunk_map[tensor]

    def get_chunks(self, tensors: Iterable[torch.Tensor]) -> Tuple[Chunk, ...]:
        return tuple(dict.fromkeys(self.tensor_chunk_map[t] for t in tensors))

    def get_cuda_movable_chunks(self) -> List[Chunk]:
        return [c for c in self.accessed_chunks if c.can_release]

    def all_chunks(self) -> List[Chunk]:
        return [c for g in self.chunk_groups.values() for c in g]

    def add_extern_static_tensor(self, tensor: torch.Tensor) -> None:
        kind = "cpu" if tensor.device.type == "cpu" else "cuda"
        self.total_mem[kind] += tensor.numel() * tensor.element_size()

    def __repr__(self) -> str:
        lines = ["ChunkManager: " + ", ".join(f"{k}={v} B" for k, v in self.total_mem.items())]
        for name, group in self.chunk_groups.items():
            lines.append(f"  group {name}: {len(group)} chunks")
            lines.extend(f"    [{i}] {c}" for i, c in enumerate(group))
        return "\n".join(lines)
