"""GeminiManager: drives chunk placement during an iteration.
Parity: reference `colossalai/zero/gemini/gemini_mgr.py:13-203`."""
from __future__ import annotations

import functools
from time import time
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from .chunk import Chunk, ChunkManager
from .memory_tracer import ChunkMemStatsCollector, MemStats
from .placement_policy import PlacementPolicy, PlacementPolicyFactory

__all__ = ["GeminiManager"]


class GeminiManager:
    def __init__(self, placement_policy: str, chunk_manager: ChunkManager, memstats: Optional[MemStats] = None,
                 max_prefetch: int = 0, **placement_kwargs) -> None:
        assert placement_policy in PlacementPolicyFactory.get_policy_names()
        self.policy_name = placement_policy
        policy_cls = PlacementPolicyFactory.create(placement_policy)
        self._chunk_manager = chunk_manager
        self._premade_memstats_ = memstats is not None
        self._memstats = memstats
        self._mem_stats_collector = ChunkMemStatsCollector(chunk_manager, self._memstats) \
            if policy_cls.need_mem_stats else None
        self._placement_policy: PlacementPolicy = policy_cls(chunk_manager, self._mem_stats_collector,
                                                             max_prefetch=max_prefetch, **placement_kwargs)
        self._compute_list: List[Tuple[Chunk, ...]] = []
        self._compute_idx: int = -1
        self._async_works: Dict[Chunk, object] = {}
        self._h2d_volume = 0
        self._d2h_volume = 0
        self._layout_time = 0.0
        self._evict_time = 0.0
        self._warmup = True
        self._comp_cuda_demand_time = 0.0

    def reset_attributes(self) -> None:
        self._compute_idx = -1
        self._h2d_volume = self._d2h_volume = 0
        self._layout_time = self._evict_time = self._comp_cuda_demand_time = 0.0

    @property
    def need_warmup(self) -> bool:
        return self.policy_name in ("auto", "const")

    def is_warmup(self) -> bool:
        return self._warmup

    def memstats(self) -> MemStats:
        return self._mem_stats_collector._memstats if self._mem_stats_collector else self._memstats

    def pre_iter(self, *args) -> None:
        if self._mem_stats_collector and self._warmup:
            self._mem_stats_collector.start_collection()

    def post_iter(self) -> None:
        if self._mem_stats_collector and self._warmup:
            self._mem_stats_collector.finish_collection()
        self._warmup = False
        self.reset_attributes()

    def adjust_layout(self, chunks: Tuple[Chunk, ...], record_anyway: bool = False) -> None:
        """Make room on the accelerator for `chunks` (evict according to the policy)."""
        start = time()
        self._record_warmup_chunks_order(chunks, record_anyway=record_anyway)
        cuda_demand, can_evict = self._get_layout_info(self._compute_idx, self._warmup, chunks)
        self._layout_time += time() - start
        vol, evict_time = self._placement_policy.evict_tensors(
            can_evict_chunks=can_evict, cuda_demand=cuda_demand, warmup=self._warmup,
            compute_list=self._compute_list, compute_idx=self._compute_idx)
        self._d2h_volume += vol
        self._evict_time += evict_time

    def wait_chunks(self, chunks: Iterable[Chunk]) -> Tuple[Chunk, ...]:
        not_prefetched = []
        for c in chunks:
            if c in self._async_works:
                w = self._async_works.pop(c)
                if w is not None:
                    w.wait()
            else:
                not_prefetched.append(c)
        return tuple(not_prefetched)

    def add_work(self, chunk: Chunk, work) -> None:
        self._async_works[chunk] = work

    def _get_layout_info(self, compute_idx: int, warmup: bool, chunks: Tuple[Chunk, ...]):
        cuda_demand = 0
        for c in chunks:
            if c.device_type == "cpu":
                cuda_demand += c.chunk_mem
            elif not c.is_gathered:
                cuda_demand += c.chunk_mem - c.shard_mem
        can_evict = [c for c in self._chunk_manager.get_cuda_movable_chunks() if c not in chunks]
        can_evict += [c for c in self._chunk_manager.all_chunks()
                      if (not c.is_gathered and c.device_type != "cpu" and c not in chunks and c not in can_evict)]
        return cuda_demand, can_evict

    def _record_warmup_chunks_order(self, chunks: Tuple[Chunk, ...], record_anyway: bool = False) -> None:
        self._compute_idx += 1
        if self._warmup and (self._placement_policy.need_mem_stats or record_anyway):
            self._compute_list.append(chunks)

    def sample_overall_data(self) -> None:
        if self._mem_stats_collector:
            self._mem_stats_collector.sample_overall_data()

    def record_model_data_volume(self) -> None:
        if self._mem_stats_collector:
            self._mem_stats_collector.record_model_data_volume()

    @property
    def chunk_manager(self) -> ChunkManager:
        return self._chunk_manager

    @property
    def cuda_margin_mem(self) -> Optional[float]:
        return self._mem_stats_collector.cuda_margin_mem if self._mem_stats_collector else None

    @property
    def placement_policy(self) -> PlacementPolicy:
        return self._placement_policy

    @property
    def compute_list(self) -> List[Tuple[Chunk, ...]]:
        return self._compute_list

    @property
    def compute_idx(self) -> int:
        return self._compute_idx

    @property
    def async_works(self) -> Dict[Chunk, object]:
        return self._async_works

    @property
    def is_cuda_margin_mem_avail(self) -> bool:
        return self._placement_policy.need_mem_stats

    def setup_grads_device(self, params: List[torch.Tensor], grads_device_map: Dict[torch.Tensor, torch.device]) -> None:
        self._placement_policy.setup_grads_device(params, grads_device_map)
