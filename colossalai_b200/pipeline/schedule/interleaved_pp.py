"""Interleaved (virtual pipeline stages) 1F1B schedule.

Parity: reference `colossalai/pipeline/schedule/interleaved_pp.py:26-616`.  Implemented on the generic node-list
executor with a node list produced by `interleaved_1f1b_schedule` (round-robin chunk placement, no dX/dW split).
"""
from __future__ import annotations

from typing import Optional

from ..stage_manager import PipelineStageManager
from .v_schedule import interleaved_1f1b_schedule
from .zero_bubble_pp import NodeListScheduler

__all__ = ["InterleavedSchedule"]


class InterleavedSchedule(NodeListScheduler):
    def __init__(self, stage_manager: PipelineStageManager, num_model_chunks: int,
                 num_microbatch: Optional[int] = None, microbatch_size: Optional[int] = None,
                 enable_metadata_cache: bool = True, overlap_p2p: bool = True, fp8_communication: bool = False) -> None:
        assert num_microbatch is not None or microbatch_size is not None, (
            "Either num_microbatch or microbatch_size should be provided")
        self._deferred = num_microbatch is None
        self._sm = stage_manager
        self._n_chunks = num_model_chunks
        sched = interleaved_1f1b_schedule(stage_manager.num_stages, num_microbatch, num_model_chunks) \
            if num_microbatch is not None else [[] for _ in range(stage_manager.num_stages)]
        super().__init__(stage_manager, sched, num_model_chunks, num_microbatch, microbatch_size, v_shape=False,
                         split_w=False, enable_metadata_cache=enable_metadata_cache, overlap_p2p=overlap_p2p)
        self.fp8_communication = fp8_communication

    def load_batch(self, data_iter, device=None) -> None:
        super().load_batch(data_iter, device)
        if self._deferred:   # number of micro-batches only known now
            self.full_schedule = interleaved_1f1b_schedule(self._sm.num_stages, self.num_microbatch, self._n_chunks)
            self._prepare(self.full_schedule)
            self._deferred = False
