"""Placement policies: where chunks live between uses.

Parity: reference `colossalai/zero/gemini/placement_policy.py:47,128` — `StaticPlacementPolicy(shard_param_frac,
offload_optim_frac, offload_param_frac)` and `AutoPlacementPolicy(warmup_non_model_data_ratio, steady_cuda_cap_ratio)`
(evict the chunks whose next use is farthest away when the accelerator budget would be exceeded).
"""
from __future__ import annotations

import functools
import warnings
from abc import ABC, abstractmethod
from time import time
from typing import Dict, List, Optional, Tuple, Type

import torch

from ...accelerator import get_accelerator
from .chunk import Chunk, ChunkManager
from .memory_tracer import ChunkMemStatsCollector

__all__ = ["PlacementPolicy", "StaticPlacementPolicy", "AutoPlacementPolicy", "PlacementPolicyFactory"]


class PlacementPolicy(ABC):
    need_mem_stats: bool = False

    def __init__(self, chunk_manager: ChunkManager, mem_stats_collector: Optional[ChunkMemStatsCollector] = None,
                 max_prefetch: int = 0, **kwargs) -> None:
        self.chunk_manager = chunk_manager
        self.mem_stats_collector = mem_stats_collector
        self.max_prefetch = max_prefetch

    @abstractmethod
    def evict_tensors(self, can_evict_chunks: List[Chunk], **kwargs) -> Tuple[int, float]:
        ...

    @abstractmethod
    def setup_grads_device(self, params: List[torch.Tensor], grads_device_map: Dict[torch.Tensor, torch.device]) -> None:
        ...

    def get_prefetch_chunks(self, is_warmup: bool, compute_list: tuple, compute_idx: int,
                            async_works: Dict[Chunk, object]) -> List[Chunk]:
        """The next `max_prefetch` distinct chunks (in recorded compute order) that are not yet on their way."""
        if is_warmup or self.max_prefetch <= 0:
            return []
        can, out = self.max_prefetch - len(async_works), []
        for i in range(compute_idx + 1, len(compute_list)):
            for chunk in compute_list[i]:
                if len(out) >= can:
                    return out
                if chunk not in out and chunk not in self.chunk_manager.accessed_chunks and chunk not in async_works:
                    out.append(chunk)
        return out


class StaticPlacementPolicy(PlacementPolicy):
    def __init__(self, chunk_manager: ChunkManager, mem_stats_collector=None, max_prefetch: int = 0,
                 shard_param_frac: float = 1.0, offload_optim_frac: float = 0.0, offload_param_frac: float = 0.0,
                 **kwargs) -> None:
        super().__init__(chunk_manager, mem_stats_collector, max_prefetch)
        if offload_param_frac > 0.0 and (shard_param_frac != 1.0 or offload_optim_frac != 1.0):
            warnings.warn("offload_param_frac is ignored when shard_param_frac != 1.0 or offload_optim_frac != 1.0")
            offload_param_frac = 0.0
        self.shard_param_frac = shard_param_frac
        self.offload_optim_frac = offload_optim_frac
        self.offload_param_frac = offload_param_frac
        self.keep_gathered_chunk_mem = 0.0
        self.keep_cuda_chunk_mem = 0.0

    def evict_tensors(self, can_evict_chunks: List[Chunk], **kwargs) -> Tuple[int, float]:
        can_shard = 0
        for c in can_evict_chunks:
            can_shard += c.chunk_mem - c.shard_mem
        # static policy: chunks are released by the DDP wrapper right after use; param offload happens here
        vol = 0
        start = time()
        if self.offload_param_frac > 0:
            for c in can_evict_chunks:
                if getattr(c, "_static_offload", False) and not c.is_gathered and c.device_type != "cpu":
                    self.chunk_manager.move_chunk(c, torch.device("cpu"))
                    vol += c.shard_mem
        return vol, time() - start

    def setup_grads_device(self, params: List[torch.Tensor], grads_device_map: Dict[torch.Tensor, torch.device]) -> None:
        """Decide per chunk: keep gathered (ZeRO-2 like) vs shard, offload optimizer shard / param shard to host."""
        total_chunk_mem = sum(self.chunk_manager.get_chunk(p).chunk_mem for p in params)
        offload_optim_mem = total_chunk_mem * self.offload_optim_frac
        offloaded = 0
        dev = get_accelerator().get_current_device()
        chunks = list(dict.fromkeys(self.chunk_manager.get_chunk(p) for p in params))
        n = len(chunks)
        n_keep = int(round(n * (1.0 - self.shard_param_frac)))
        n_off_param = int(round(n * self.offload_param_frac))
        for i, c in enumerate(chunks):
            c._static_keep_gathered = i < n_keep
            c._static_offload = i >= n - n_off_param
        for p in params:
            c = self.chunk_manager.get_chunk(p)
            if offloaded < offload_optim_mem:
                grads_device_map[p] = torch.device("cpu")
                if getattr(c, "_counted_optim", False) is False:
                    offloaded += c.chunk_mem
                    c._counted_optim = True
            else:
                grads_device_map[p] = dev


class AutoPlacementPolicy(PlacementPolicy):
    need_mem_stats: bool = True

    def __init__(self, chunk_manager: ChunkManager, mem_stats_collector: Optional[ChunkMemStatsCollector] = None,
                 max_prefetch: int = 0, warmup_non_model_data_ratio: float = 0.8, steady_cuda_cap_ratio: float = 0.9,
                 **kwargs) -> None:
        super().__init__(chunk_manager, mem_stats_collector, max_prefetch)
        self._warmup_non_model_data_ratio = warmup_non_model_data_ratio
        self._steady_cuda_cap_ratio = steady_cuda_cap_ratio

    def evict_tensors(self, can_evict_chunks: List[Chunk], cuda_demand: int = 0, warmup: bool = True,
                      compute_list: Optional[List[Tuple[Chunk, ...]]] = None, compute_idx: int = 0,
                      **kwargs) -> Tuple[int, float]:
        start = time()
        cuda_capacity = get_accelerator().mem_get_info()[1] if torch.cuda.is_available() else 1 << 62
        used_cuda_model_data = self.chunk_manager.total_mem["cuda"]
        if warmup:
            max_nmd = cuda_capacity * self._warmup_non_model_data_ratio
        else:
            max_nmd = self.mem_stats_collector.next_period_non_model_data_usage("cuda")
            cuda_capacity *= self._steady_cuda_cap_ratio
        total_model_data = cuda_capacity - max_nmd
        avail = total_model_data - used_cuda_model_data
        freed = 0
        if avail < cuda_demand:
            to_free = cuda_demand - avail
            chunks = can_evict_chunks
            if not warmup and compute_list is not None:
                chunks = self._sort_can_evict_chunks(tuple(chunks), compute_idx, tuple(compute_list))
            for c in chunks:
                if freed >= to_free:
                    break
                self.chunk_manager.release_chunk(c)
                self.chunk_manager.move_chunk(c, torch.device("cpu"))
                freed += c.chunk_mem
            if freed < to_free:
                raise RuntimeError(f"Adjust layout failed! No enough CUDA memory! Need {to_free}, freed {freed}")
        return freed, time() - start

    @staticmethod
    @functools.lru_cache(maxsize=None)
    def _sort_can_evict_chunks(can_evict_chunks: tuple, compute_idx: int, compute_list: tuple) -> list:
        next_use = {c: len(compute_list) for c in can_evict_chunks}
        for i in range(len(compute_list) - 1, compute_idx, -1):
            for c in compute_list[i]:
                if c in next_use:
                    next_use[c] = i
        return [c for c, _ in sorted(next_use.items(), key=lambda kv: kv[1], reverse=True)]

    def setup_grads_device(self, params: List[torch.Tensor], grads_device_map: Dict[torch.Tensor, torch.device]) -> None:
        dev = get_accelerator().get_current_device()
        for p in params:
            c = self.chunk_manager.get_chunk(p)
            grads_device_map[p] = dev if c.keep_gathered else torch.device("cpu")


class PlacementPolicyFactory:
    policies: Dict[str, Type[PlacementPolicy]] = {"auto": AutoPlacementPolicy, "static": StaticPlacementPolicy}

    @staticmethod
    def create(policy_name: str) -> Type[PlacementPolicy]:
        if policy_name not in PlacementPolicyFactory.policies:
            raise TypeError(f"Unknown tensor placement policy {policy_name}")
        return PlacementPolicyFactory.policies[policy_name]

    @staticmethod
    def get_policy_names():
        return tuple(PlacementPolicyFactory.policies.keys())
